"""
ORACLE (test infrastructure) - CPU restatement of the foldingdiff noise predictor.

Follows, op for op and in the reference's evaluation order:
  * foldingdiff/modelling.py:384-484  BertForDiffusionBase.forward
  * foldingdiff/modelling.py:59-71    GaussianFourierProjection.forward
  * foldingdiff/modelling.py:157-170  BertEmbeddings.forward (relative_key: no abs. pos. emb.)
  * foldingdiff/modelling.py:203-208  AnglesPredictor.forward
  * transformers==4.11.3 models/bert/modeling_bert.py BertEncoder / BertLayer /
    BertSelfAttention(position_embedding_type="relative_key") / BertSelfOutput /
    BertIntermediate / BertOutput (call site modelling.py:473-480) - restated from
    the published algorithm because that wheel is not installable here
    (**parity unpinned** at this boundary, see oracle/__init__.py).

Eval mode (dropout = identity).  Works in fp32 (default, = the reference's
arithmetic) or fp64 (for error budgeting).  Pure torch CPU ops.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


class OracleConfig:
    """The subset of HF BertConfig the forward reads (config.json keys)."""

    def __init__(self, hidden_size, num_hidden_layers, num_attention_heads,
                 intermediate_size, max_position_embeddings=128,
                 layer_norm_eps=1e-12, position_embedding_type="relative_key",
                 **_unused):
        self.hidden_size = int(hidden_size)
        self.num_hidden_layers = int(num_hidden_layers)
        self.num_attention_heads = int(num_attention_heads)
        self.intermediate_size = int(intermediate_size)
        self.max_position_embeddings = int(max_position_embeddings)
        self.layer_norm_eps = float(layer_norm_eps)
        self.position_embedding_type = position_embedding_type


def _lin(sd: Dict[str, torch.Tensor], key: str, z: torch.Tensor) -> torch.Tensor:
    return F.linear(z, sd[key + ".weight"], sd[key + ".bias"])


def _ln(sd, key, z, eps):
    return F.layer_norm(z, (z.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def time_embedding(W: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """modelling.py:69-70 - x[:, None] * W[None, :] * 2 * pi, left to right, in W's dtype."""
    x = t.to(W.dtype)
    x_proj = x[:, None] * W[None, :] * 2 * torch.pi
    return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, x: torch.Tensor,
            t: torch.Tensor, attention_mask: torch.Tensor,
            time_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """
    x (B,N,F), t (B,) int64, attention_mask (B,N) in {0,1}  ->  eps_hat (B,N,F).
    `sd` is the checkpoint state_dict (reference key names), already in the
    dtype to compute in.  `time_table` (T,H), if given, replaces the sin/cos
    evaluation (used to run the fp64 oracle with the fp32 time embedding).
    """
    B, N, _ = x.shape
    H, nh = cfg.hidden_size, cfg.num_attention_heads
    dh = H // nh
    eps = cfg.layer_norm_eps
    assert attention_mask.dim() == 2  # modelling.py:447-449
    ext = (1.0 - attention_mask.to(x.dtype))[:, None, None, :] * -10000.0  # :450-452

    h = _lin(sd, "inputs_to_hidden_dim", x)  # :464
    h = _ln(sd, "embeddings.LayerNorm", h, eps)  # :168 (relative_key: no position table)
    if time_table is not None:
        te = time_table[t]
    else:
        te = time_embedding(sd["time_embed.W"], t)  # :471
    h = h + te.to(h.dtype)[:, None, :]  # :472

    pos = torch.arange(N)
    dist = pos[:, None] - pos[None, :] + (cfg.max_position_embeddings - 1)
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{l}."
        q = _lin(sd, p + "attention.self.query", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
        k = _lin(sd, p + "attention.self.key", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
        v = _lin(sd, p + "attention.self.value", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
        s = torch.matmul(q, k.transpose(-1, -2))
        if cfg.position_embedding_type == "relative_key":
            E = sd[p + "attention.self.distance_embedding.weight"][dist]  # (N,N,dh)
            s = s + torch.einsum("bhld,lrd->bhlr", q, E)
        elif cfg.position_embedding_type != "absolute":
            raise NotImplementedError(cfg.position_embedding_type)
        s = s / math.sqrt(dh)
        s = s + ext
        pr = torch.softmax(s, dim=-1)
        c = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, N, H)
        a = _ln(sd, p + "attention.output.LayerNorm", _lin(sd, p + "attention.output.dense", c) + h, eps)
        i = F.gelu(_lin(sd, p + "intermediate.dense", a))  # HF "gelu" = exact erf
        h = _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", i) + a, eps)

    d = _lin(sd, "token_decoder.dense1", h)  # modelling.py:204
    d = F.gelu(d)  # :205
    d = _ln(sd, "token_decoder.layer_norm", d, 1e-12)  # :206 (AnglesPredictor eps default)
    return _lin(sd, "token_decoder.dense2", d)  # :207


class OracleModel(torch.nn.Module):
    """
    nn.Module wrapper so that the reference's own sampling.p_sample /
    p_sample_loop (which call ``model(x, t, attention_mask=...)`` and
    ``next(model.parameters()).device``) can drive the oracle unmodified.
    """

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: OracleConfig,
                 ft_is_angular, dtype=torch.float32):
        super().__init__()
        self.cfg = cfg
        self.config = cfg
        self.ft_is_angular = list(ft_is_angular)
        self.n_inputs = len(self.ft_is_angular)
        self.dtype_ = dtype
        self._keys = list(sd.keys())
        for i, k in enumerate(self._keys):
            self.register_parameter(
                f"p{i}", torch.nn.Parameter(sd[k].detach().to(dtype).clone(), requires_grad=False))
        self._time_table = None

    def set_time_table(self, table: Optional[torch.Tensor]):
        self._time_table = table

    def state(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, f"p{i}") for i, k in enumerate(self._keys)}

    def forward(self, inputs, timestep, attention_mask, position_ids=None):
        out = forward(self.state(), self.cfg, inputs.to(self.dtype_), timestep,
                      attention_mask, time_table=self._time_table)
        return out
