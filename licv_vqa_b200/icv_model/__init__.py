from .icv_intervention import LearnableICVInterventionLMM

__all__ = ["LearnableICVInterventionLMM"]
