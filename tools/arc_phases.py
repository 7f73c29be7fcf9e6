"""Development aid: per-launch CUDA-event times of one eager ArcFace step (DIF_ARC_PROFILE=1)."""
import os
import sys

os.environ["DIF_ARC_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deep_insight_face_b200.arcface import ArcFaceStep

B, C, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 10000, 512)
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
st = ArcFaceStep(B, C, D, graph=False, precision=prec)
st.y.copy_(torch.randint(0, C, (B,), device="cuda").int())
for _ in range(8):
    st()
torch.cuda.synchronize()
