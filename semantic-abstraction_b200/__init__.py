"""semantic-abstraction_b200 — B200-native hot path of real-stanford/semantic-abstraction.

Host-side Python mirror of the reference API (ClipWrapper.get_clip_saliency, SemAbs3D.forward, ResidualUNet3D)
over hand-written sm_100a CUDA kernels behind the C ABI in include/semabs_b200.h. Import as `semabs_b200`.
"""
__version__ = "0.1.0"
