"""raw2film_b200: B200-native (sm_100a) implementation of raw2film's per-pixel film-emulation
render path behind the reference's processor API.  See DESIGN.md."""
from __future__ import annotations

__all__ = ["B200Processor", "SyntheticStock", "BatchExporter", "PipelinedRenderer", "PreviewGraph"]


def __getattr__(name):  # lazy: importing the package does not need CUDA, using the processor does
    if name == "B200Processor":
        from .processor import B200Processor

        return B200Processor
    if name == "SyntheticStock":
        from .synthetic import SyntheticStock

        return SyntheticStock
    if name == "BatchExporter":
        from .batch import BatchExporter

        return BatchExporter
    if name == "PipelinedRenderer":
        from .pipeline import PipelinedRenderer

        return PipelinedRenderer
    if name == "PreviewGraph":
        from .pipeline import PreviewGraph

        return PreviewGraph
    raise AttributeError(name)
