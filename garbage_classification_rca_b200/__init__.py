"""B200-native MM-RCA late-fusion head (drop-in for the hot path of espiriki/Garbage_Classification_RCA).

Host side: Python/PyTorch mirror of the reference model API (multimodal_model.py).
Device side: hand-written sm_100a CUDA kernels behind a C ABI (include/mmrca.h, libmmrca.so).
"""
from . import _native, functional
from .functional import (FusionTrainStep, HeadTrainStep, HierTrainStep, attention_block, concat_width, cross_entropy,
                         feature_handoff, final_linear_name, fusion_head, head_param_names, hierarchical_head, mmrca_head)

__all__ = ["_native", "functional", "FusionTrainStep", "HeadTrainStep", "HierTrainStep", "attention_block", "concat_width",
           "cross_entropy", "feature_handoff", "final_linear_name", "fusion_head", "head_param_names", "hierarchical_head",
           "mmrca_head"]
