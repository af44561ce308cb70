"""
Seeded synthetic weights and inputs (bench + test support; NOT part of oracle/).

The real production weights (wukevin/foldingdiff_cath on the HF hub) are not
available offline, so BASELINE configs 2-5 run on random weights of exactly
that architecture (config_jsons/cath_full_angles_cosine.json:3-14), following
SURVEY.md section 8(d) "Config 2": Linear/embedding weights and biases N(0, 0.02),
LayerNorm gamma = 1 + N(0, 0.1), beta = N(0, 0.1), distance_embedding N(0, 0.1),
time_embed.W = randn(H/2) * 2*pi, all from torch.Generator().manual_seed(seed).
State-dict key names are the reference checkpoint's.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

PRODUCTION = dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=12,
                  intermediate_size=768, max_position_embeddings=128,
                  layer_norm_eps=1e-12, position_embedding_type="relative_key")
MINI = dict(hidden_size=192, num_hidden_layers=6, num_attention_heads=6,
            intermediate_size=384, max_position_embeddings=128,
            layer_norm_eps=1e-12, position_embedding_type="relative_key")
# jupyter/test_set_partial_denoise.ipynb:131 (training-set mean offsets of the CATH model)
CATH_MEAN_OFFSET = np.array([-1.4702034, 0.0361131, 3.1276708, 1.9405054, 2.0354161, 2.1225433],
                            dtype=np.float32)


def synthetic_state_dict(cfg: dict, n_features: int = 6, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    dh = H // cfg["num_attention_heads"]
    P = cfg["max_position_embeddings"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * 0.02
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def ln(name, width):
        sd[name + ".weight"] = 1.0 + torch.randn(width, generator=g) * 0.1
        sd[name + ".bias"] = torch.randn(width, generator=g) * 0.1

    lin("inputs_to_hidden_dim", H, n_features)
    ln("embeddings.LayerNorm", H)
    for l in range(L):
        p = f"encoder.layer.{l}."
        lin(p + "attention.self.query", H, H)
        lin(p + "attention.self.key", H, H)
        lin(p + "attention.self.value", H, H)
        sd[p + "attention.self.distance_embedding.weight"] = torch.randn(2 * P - 1, dh, generator=g) * 0.1
        lin(p + "attention.output.dense", H, H)
        ln(p + "attention.output.LayerNorm", H)
        lin(p + "intermediate.dense", I, H)
        lin(p + "output.dense", H, I)
        ln(p + "output.LayerNorm", H)
    lin("token_decoder.dense1", H, H)
    ln("token_decoder.layer_norm", H)
    lin("token_decoder.dense2", n_features, H)
    sd["time_embed.W"] = torch.randn(H // 2, generator=g) * 2 * torch.pi
    return sd


def sweep_lengths(batch: int, lo: int = 50, hi: int = 128) -> List[int]:
    """lengths[i] = lo + (i mod (hi-lo)): the 50..127 mix of sampling.py:169-170."""
    return [lo + (i % (hi - lo)) for i in range(batch)]


def algorithmic_flops(cfg: dict, lengths, n_features: int = 6) -> float:
    """F(n) = F_tok*n + F_att*n^2 per chain per reverse step."""
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    f_tok = 2 * n_features * H + L * (8 * H * H + 4 * H * I) + 2 * H * H + 2 * H * n_features
    f_att = L * 6 * H
    n = np.asarray(lengths, dtype=np.float64)
    return float(f_tok * n.sum() + f_att * (n * n).sum())
