"""Import path of the reference (icv_src/icv_encoder/global_icv_encoder.py); the code is in encoders.py."""
from .encoders import GlobalICVEncoder  # noqa: F401
