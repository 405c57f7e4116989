"""infinisst_b200: B200-native per-chunk streaming step of InfiniSST."""
from .config import (EncoderConfig, LlmConfig, TemplateConfig, GenConfig, InfiniSSTConfig,
                     production_config, tiny_config)

__all__ = ["EncoderConfig", "LlmConfig", "TemplateConfig", "GenConfig", "InfiniSSTConfig",
           "production_config", "tiny_config"]
