"""
Cross-check of the restated encoder layer against HuggingFace's OWN modules (the installed transformers, not 4.11.3).

The reference's encoder arithmetic lives in transformers==4.11.3 (`position_embedding_type="relative_key"`), which
cannot be installed here, and the installed transformers' BERT no longer has that branch - so the oracle restates it
(oracle/forward.py).  What CAN be checked against HF code:
  * `Wav2Vec2BertSelfAttention(position_embeddings_type="relative_key")` is HF's surviving implementation of the same
    scheme: scores = q k^T / sqrt(d) + einsum("bhld,lrd->bhlr", q, E[distance]) / sqrt(d) + mask, softmax, P v, output
    projection.  It indexes the table with r - l where BERT 4.11.3 uses l - r, i.e. with the table reversed
    (E_w2v[j] = E_bert[254 - j] for max_position_embeddings = 128) it is BertSelfAttention + BertSelfOutput.dense.
  * `BertSelfOutput`, `BertIntermediate`, `BertOutput` of the installed transformers are the residual / LayerNorm /
    GELU blocks the reference's BertLayer is made of.
Composed, they must reproduce the oracle's layer on the real mini-fixture weights.  This pins the einsum layout, the
scaling, the additive mask, the softmax axis, the head split and the sub-block order to HF's code; the direction of the
distance index (l - r) remains the one statement taken from the 4.11.3 source text.
"""
import types

import pytest
import torch

from conftest import mini_state_dict
from oracle import forward as ofwd

hf_bert = pytest.importorskip("transformers.models.bert.modeling_bert")
hf_w2v = pytest.importorskip("transformers.models.wav2vec2_bert.modeling_wav2vec2_bert")


def _hf_layer(sd, cfg, layer):
    H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
    P = cfg.get("max_position_embeddings", 128)
    p = f"encoder.layer.{layer}."
    acfg = types.SimpleNamespace(hidden_size=H, num_attention_heads=nh, position_embeddings_type="relative_key",
                                 left_max_position_embeddings=P - 1, right_max_position_embeddings=P - 1, attention_dropout=0.0)
    att = hf_w2v.Wav2Vec2BertSelfAttention(acfg)
    with torch.no_grad():
        for mine, theirs in (("linear_q", "attention.self.query"), ("linear_k", "attention.self.key"),
                             ("linear_v", "attention.self.value"), ("linear_out", "attention.output.dense")):
            getattr(att, mine).weight.copy_(sd[p + theirs + ".weight"])
            getattr(att, mine).bias.copy_(sd[p + theirs + ".bias"])
        att.distance_embedding.weight.copy_(sd[p + "attention.self.distance_embedding.weight"].flip(0))  # r - l  <->  l - r
    from transformers import BertConfig
    bcfg = BertConfig(hidden_size=H, num_attention_heads=nh, intermediate_size=cfg["intermediate_size"], hidden_act="gelu",
                      layer_norm_eps=cfg.get("layer_norm_eps", 1e-12), hidden_dropout_prob=0.0)
    self_out, inter, out = hf_bert.BertSelfOutput(bcfg), hf_bert.BertIntermediate(bcfg), hf_bert.BertOutput(bcfg)
    with torch.no_grad():
        self_out.dense.weight.copy_(torch.eye(H)); self_out.dense.bias.zero_()  # the projection already ran inside `att`
        self_out.LayerNorm.weight.copy_(sd[p + "attention.output.LayerNorm.weight"]); self_out.LayerNorm.bias.copy_(sd[p + "attention.output.LayerNorm.bias"])
        inter.dense.weight.copy_(sd[p + "intermediate.dense.weight"]); inter.dense.bias.copy_(sd[p + "intermediate.dense.bias"])
        out.dense.weight.copy_(sd[p + "output.dense.weight"]); out.dense.bias.copy_(sd[p + "output.dense.bias"])
        out.LayerNorm.weight.copy_(sd[p + "output.LayerNorm.weight"]); out.LayerNorm.bias.copy_(sd[p + "output.LayerNorm.bias"])
    for m in (att, self_out, inter, out):
        m.eval()

    @torch.no_grad()
    def layer_fn(h, ext_mask):
        ctx, _ = att(h, attention_mask=ext_mask)
        a = self_out(ctx, h)
        return out(inter(a), a)
    return layer_fn


def _oracle_layer(sd, ocfg, layer, h, mask01):
    """One encoder layer of oracle.forward, by running the restated forward on a 1-layer view of the weights."""
    B, N, H = h.shape
    nh, dh, eps = ocfg.num_attention_heads, H // ocfg.num_attention_heads, ocfg.layer_norm_eps
    p = f"encoder.layer.{layer}."
    ext = (1.0 - mask01)[:, None, None, :] * -10000.0
    pos = torch.arange(N)
    dist = pos[:, None] - pos[None, :] + (ocfg.max_position_embeddings - 1)
    q = ofwd._lin(sd, p + "attention.self.query", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
    k = ofwd._lin(sd, p + "attention.self.key", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
    v = ofwd._lin(sd, p + "attention.self.value", h).view(B, N, nh, dh).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) + torch.einsum("bhld,lrd->bhlr", q, sd[p + "attention.self.distance_embedding.weight"][dist])
    pr = torch.softmax(s / dh ** 0.5 + ext, dim=-1)
    c = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, N, H)
    a = ofwd._ln(sd, p + "attention.output.LayerNorm", ofwd._lin(sd, p + "attention.output.dense", c) + h, eps)
    i = torch.nn.functional.gelu(ofwd._lin(sd, p + "intermediate.dense", a))
    return ofwd._ln(sd, p + "output.LayerNorm", ofwd._lin(sd, p + "output.dense", i) + a, eps)


@pytest.mark.parametrize("layer", [0, 3, 5])
def test_oracle_layer_equals_composition_of_hf_modules(layer):
    sd, cfg, _, _ = mini_state_dict()
    ocfg = ofwd.OracleConfig(**cfg)
    g = torch.Generator().manual_seed(100 + layer)
    B, N = 3, 128
    h = torch.randn(B, N, cfg["hidden_size"], generator=g)
    lengths = [128, 77, 50]
    mask = torch.zeros(B, N)
    for b, n in enumerate(lengths):
        mask[b, :n] = 1.0
    ext = (1.0 - mask)[:, None, None, :] * -10000.0
    got = _hf_layer(sd, cfg, layer)(h, ext)
    want = _oracle_layer(sd, ocfg, layer, h, mask)
    valid = mask.bool()
    err = float((got - want)[valid].abs().max())
    print(f"layer {layer}: max |HF composition - oracle| over valid rows = {err:.3e}")
    assert err < 5e-6
    # the relative-key term matters at this tolerance: with the table NOT reversed the same composition is far off
    sd_wrong = dict(sd)
    key = f"encoder.layer.{layer}.attention.self.distance_embedding.weight"
    sd_wrong[key] = sd[key].flip(0)
    off = float((_hf_layer(sd_wrong, cfg, layer)(h, ext) - want)[valid].abs().max())
    assert off > 100 * err


def test_whole_oracle_forward_is_that_layer_stacked():
    """oracle.forward == embeddings + the layer above applied num_hidden_layers times + decoder (guards the helper)."""
    sd, cfg, _, _ = mini_state_dict()
    ocfg = ofwd.OracleConfig(**cfg)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 64, 6, generator=g)
    t = torch.tensor([3, 200])
    mask = torch.ones(2, 64); mask[1, 40:] = 0
    ref = ofwd.forward(sd, ocfg, x, t, mask)
    h = ofwd._ln(sd, "embeddings.LayerNorm", ofwd._lin(sd, "inputs_to_hidden_dim", x), ocfg.layer_norm_eps)
    h = h + ofwd.time_embedding(sd["time_embed.W"], t)[:, None, :]
    for l in range(ocfg.num_hidden_layers):
        h = _oracle_layer(sd, ocfg, l, h, mask)
    d = ofwd._ln(sd, "token_decoder.layer_norm", torch.nn.functional.gelu(ofwd._lin(sd, "token_decoder.dense1", h)), 1e-12)
    assert torch.equal(ofwd._lin(sd, "token_decoder.dense2", d), ref)
