from .streaming_pipeline import LocalWhisperBackend, StreamingPipeline, StreamScheduler, TranscriptionBackend

__all__ = ["StreamingPipeline", "StreamScheduler", "TranscriptionBackend", "LocalWhisperBackend"]
