"""B200-native Semantic-NeRF rendering engine: drop-in for the nr4seg/nerf hot path of
ethz-asl/ucsa_neural_rendering.  CUDA kernels (sm_100a) live in csrc/ behind the C ABI of
include/ucsa_nerf.h; ``nerf`` mirrors the reference's Python interface on top of them."""

__version__ = "0.1.0"
