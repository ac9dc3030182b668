"""Drop-in for the reference's 3_Models/backbones/art.py: re-exports the B200 implementation under the module name the
reference's scripts import by file path (4_Experiments/scripts/train_art.py:31-44, train_multimodal_fuzzy_fusion.py:62-88).
Copy / symlink overlay/3_Models over the reference's 3_Models."""
from eyegaze_multimodal_b200.art import *  # noqa: F401,F403
from eyegaze_multimodal_b200 import art as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
