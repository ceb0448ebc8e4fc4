from .encoders import BaseICVEncoder, GlobalICVEncoder, ICVEncoderOutput

__all__ = ["BaseICVEncoder", "ICVEncoderOutput", "GlobalICVEncoder"]
