"""Drop-in for the reference's models/sit.py: `from models.sit import SiT` (tools/train.py:36, tools/testing.py)."""
from surface_vision_transformers_b200.sit import SiT  # noqa: F401
