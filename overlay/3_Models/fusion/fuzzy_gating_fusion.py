"""Drop-in for the reference's 3_Models/fusion/fuzzy_gating_fusion.py."""
from eyegaze_multimodal_b200.fuzzy_gating_fusion import *  # noqa: F401,F403
from eyegaze_multimodal_b200 import fuzzy_gating_fusion as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
