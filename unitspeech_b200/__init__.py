"""unitspeech_b200: B200-native (sm_100a) reverse-diffusion mel decoder with the call surface of UnitSpeech,
plus the BigVGAN vocoder stage that follows it."""

from .decoder import GradLogPEstimator2d, UnitSpeech, denormalize_mel, normalize_mel  # noqa: F401
from .checkpoint import DecoderBundle, load_decoder_checkpoint, save_decoder_checkpoint  # noqa: F401
from .vocoder import AttrDict, BigVGAN, get_vocoder  # noqa: F401
from .training import FineTuner  # noqa: F401
from .util import fix_len_compatibility, generate_path, sequence_mask  # noqa: F401

__all__ = ["UnitSpeech", "GradLogPEstimator2d", "denormalize_mel", "normalize_mel", "fix_len_compatibility",
           "generate_path", "sequence_mask", "DecoderBundle", "load_decoder_checkpoint", "save_decoder_checkpoint", "BigVGAN", "AttrDict", "get_vocoder", "FineTuner"]
