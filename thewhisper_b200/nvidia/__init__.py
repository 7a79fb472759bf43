from .asr_pipeline import ASRPipeline

__all__ = ["ASRPipeline"]
