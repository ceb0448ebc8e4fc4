"""Import path of the reference (icv_src/icv_encoder/base_icv_encoder.py); the code is in encoders.py."""
from .encoders import BaseICVEncoder, ICVEncoderOutput  # noqa: F401
