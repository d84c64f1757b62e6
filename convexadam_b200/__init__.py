"""convexadam_b200 -- B200-native implementation of the convexAdam dense 3-D registration hot path.

Entry points mirror the reference package (multimodallearning/convexAdam, src/convexAdam):
    convexadam_b200.convex_adam_MIND.convex_adam_pt / convex_adam / extract_features
    convexadam_b200.convex_adam_nnUNet.convex_adam / extract_features
    convexadam_b200.convex_adam_utils.MINDSSC / correlate / coupled_convex / inverse_consistency
All compute runs in hand-written sm_100a kernels (convexadam_b200/csrc) behind the C ABI of
include/convexadam_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
