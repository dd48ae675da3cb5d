"""DynChunkTrainConfig -- mirror of speechbrain.utils.dynamic_chunk_training.DynChunkTrainConfig
(utils/dynamic_chunk_training.py:24-58): the run-time configuration of chunked ("streaming-equivalent") evaluation that
``TransformerASR.encode(..., dynchunktrain_config=...)`` takes.  The random sampler used at training time is not mirrored."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class DynChunkTrainConfig:
    chunk_size: int
    """Size in frames of a single chunk, always > 0."""

    left_context_size: Optional[int] = None
    """Number of *chunks* (not frames) visible to the left, >= 0; None = infinite left context."""

    def is_infinite_left_context(self) -> bool:
        return self.left_context_size is None

    def left_context_size_frames(self) -> Optional[int]:
        if self.left_context_size is None:
            return None
        return self.chunk_size * self.left_context_size
