"""Time the CUDA-graphed ArcFace step (C2: B=512, C=10000, D=512) - development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_insight_face_b200.arcface import ArcFaceStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
st = ArcFaceStep(B, C, D, graph=True, precision=prec)
st.y.copy_(torch.randint(0, C, (B,), device="cuda").int())
for _ in range(5):
    st()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    st()
e1.record()
torch.cuda.synchronize()
print("arcface %s B=%d C=%d D=%d us/step %.1f" % (prec, B, C, D, e0.elapsed_time(e1) * 10))
