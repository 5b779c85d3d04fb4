from .config import ModelConfig, get_model_config, PRESETS
from .transformer import (PackedInput, DecodeInput, TransformerWeights, forward_hidden, layer_forward, layer_decode,
                          embed_forward, final_norm, logits_forward, value_forward, lm_head_weight,
                          layer_param_specs, embed_param_specs, head_param_specs, init_param_, pack_padded,
                          unpack_to_padded, convert_hf_state_dict, export_hf_state_dict)
