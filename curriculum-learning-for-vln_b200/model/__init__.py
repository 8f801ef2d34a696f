"""nn.Modules of the three agents (reference: src/model/units.py, policy.py) on the CUDA kernels."""
from .units import (EncoderLSTM, SoftDotAttention, VisualSoftDotAttention, ActionScoring, PositionalEncoding,
                    MLPwithBN, LengthMask, use_rng)
from .policy import AttnDecoderLSTM, MonitorDecoder, EnvDropDecoder, Critic
from .speaker import SpeakerEncoder, SpeakerDecoder

__all__ = ["EncoderLSTM", "SoftDotAttention", "VisualSoftDotAttention", "ActionScoring", "PositionalEncoding",
           "MLPwithBN", "LengthMask", "use_rng", "AttnDecoderLSTM", "MonitorDecoder", "EnvDropDecoder", "Critic",
           "SpeakerEncoder", "SpeakerDecoder"]
