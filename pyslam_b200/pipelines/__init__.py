"""Callers of the hot path (reference pyslam/pipelines/): RANSAC, keyframe pyramids and the dense VO loop,
with their data-parallel parts on the GPU."""
from .dense import DenseRGBDPipeline, DenseStereoPipeline, DenseVOPipeline
from .keyframes import (DenseKeyframe, DenseRGBDKeyframe, DenseStereoKeyframe, Keyframe, SparseRGBDKeyframe,
                        SparseStereoKeyframe)
from .ransac import FrameToFrameRANSAC
from .sparse import SparseRGBDPipeline, SparseStereoPipeline, SparseVOPipeline

__all__ = ['FrameToFrameRANSAC', 'Keyframe', 'DenseKeyframe', 'DenseRGBDKeyframe', 'DenseStereoKeyframe',
           'SparseStereoKeyframe', 'SparseRGBDKeyframe', 'DenseVOPipeline', 'DenseStereoPipeline', 'DenseRGBDPipeline',
           'SparseVOPipeline', 'SparseStereoPipeline', 'SparseRGBDPipeline']
