"""deep_insight_face_b200 - the embedding-space distance path of sandyz1000/deep-insight-face on NVIDIA B200.

Hand-written sm_100a kernels (tcgen05 / TMEM / TMA) behind a C ABI (include/dif_b200.h), bound with
ctypes (_ffi.py) and exposed through the reference's own names:

    common.losses        BatchHardTripletLoss, BatchHardTripletLossEuclidean, ...AutoAlpha, BatchAllTripletLoss
    common.tfa_losses    TripletHardLoss, TripletSemiHardLoss (the tfa.losses the reference compiles with)
    networks.triplet     triplet_loss
    networks.siamese     euclidean_distance, contrastive_loss
    networks.head        l2_normalize (the embedding head's Lambda)
    networks.utils       distance, distance_to_proba, gaussian_kernel_dist_to_prob
    evaluation.utility   distance, calculate_accuracy, calculate_val_far, calculate_roc, calculate_val, evaluate
    api                  face_distance, compare_faces
    predictions          TripletPrediction.verify, SiamesePrediction.verify
    gallery              Gallery, ShardedGallery (1:N top-k search, new)
    arcface              arcface_loss, ArcFaceLoss (new)
    datagen              sample_people, pk_labels, create_pairs, facematch_image_pairs, triplet_image_pairs,
                         write_pairs_to_file (host-side callers of the losses)

There is no CPU fallback: importing is cheap, but any compute call needs libdif_b200.so and a B200.
"""
__version__ = "0.1.0"
