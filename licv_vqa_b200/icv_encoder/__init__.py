from .base_icv_encoder import BaseICVEncoder, ICVEncoderOutput
from .global_icv_encoder import GlobalICVEncoder

__all__ = ["BaseICVEncoder", "ICVEncoderOutput", "GlobalICVEncoder"]
