"""
`bench.py --impl reference`: the reference's own CPU path, as far as this image lets it run.

baseline/_ref holds the UNMODIFIED reference package, installed by
    cp -r /root/reference /tmp/refcopy && python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref /tmp/refcopy
(`--no-deps`: pytorch_lightning, biotite and transformers==4.11.3 are not in the wheelhouse; the source tree is read-only,
hence the copy).  baseline/_ref is git-ignored but travels to the GPU box.  What runs from it, stock:
  * `foldingdiff.sampling.p_sample_loop` / `p_sample`  - the hot loop itself, with its per-chain mask loop, its
    per-step `compute_alphas`, its `.item()` syncs (sampling.py:28-132)
  * `foldingdiff.beta_schedules`, `foldingdiff.utils.modulo_with_wrapped_range`
  * `foldingdiff.modelling.GaussianFourierProjection`, `BertEmbeddings`, `AnglesPredictor`  - the model's own sub-modules
The import needs stubs for matplotlib / pytorch_lightning / biotite (oracle/ref_shims.py: no arithmetic lives there).

What cannot run: `BertForDiffusionBase.__init__` (modelling.py:291 -> `init_weights()` -> AttributeError
`all_tied_weights_keys` on transformers 5.5) and the HF 4.11.3 `BertEncoder` with `relative_key` (gone from BERT in 5.5).
The encoder layers are therefore assembled from the INSTALLED transformers' own modules - the same third-party library the
reference delegates to: `Wav2Vec2BertSelfAttention(position_embeddings_type="relative_key")` (HF's surviving
implementation of the scheme; distance r - l, so the table is loaded reversed) + `BertSelfOutput` + `BertIntermediate` +
`BertOutput`, wired in the order of modelling.py:427-484.  tests/test_reference_arm.py checks this module against the
oracle forward (1e-6) where baseline/_ref exists.  None of this repo's kernels, engine or models is on this path.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "foldingdiff", "sampling.py"))


def load_reference():
    """-> (sampling, beta_schedules, utils, modelling) modules of the installed reference."""
    os.environ["FOLDINGDIFF_REFERENCE"] = REF_DIR
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_shims  # import-time stubs only
    ref_shims.REFERENCE_ROOT = REF_DIR
    ref_shims.install()
    loaded = sys.modules.get("foldingdiff")
    if loaded is not None and not os.path.realpath(getattr(loaded, "__file__", "") or "").startswith(os.path.realpath(REF_DIR)):
        # the authoring container's tests may already have imported the same package from /root/reference
        for name in [n for n in sys.modules if n == "foldingdiff" or n.startswith("foldingdiff.")]:
            del sys.modules[name]
    if REF_DIR in sys.path:
        sys.path.remove(REF_DIR)
    sys.path.insert(0, REF_DIR)
    from foldingdiff import beta_schedules, modelling, sampling, utils  # type: ignore
    assert os.path.realpath(sampling.__file__).startswith(os.path.realpath(REF_DIR)), sampling.__file__
    return sampling, beta_schedules, utils, modelling


def build_model(state_dict, cfg: dict):
    """The reference's forward (modelling.py:427-484) from its own sub-modules + HF library blocks; eval mode, CPU."""
    import torch
    from torch import nn
    from transformers import BertConfig
    from transformers.models.bert import modeling_bert as hf_bert
    from transformers.models.wav2vec2_bert import modeling_wav2vec2_bert as hf_w2v

    _, _, _, rm = load_reference()
    H, nh, L = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_hidden_layers"]
    P = cfg.get("max_position_embeddings", 128)
    n_in = state_dict["inputs_to_hidden_dim.weight"].shape[1]
    bcfg = BertConfig(hidden_size=H, num_hidden_layers=L, num_attention_heads=nh, intermediate_size=cfg["intermediate_size"],
                      max_position_embeddings=P, position_embedding_type="relative_key", hidden_act="gelu",
                      layer_norm_eps=cfg.get("layer_norm_eps", 1e-12), hidden_dropout_prob=cfg.get("hidden_dropout_prob", 0.1),
                      attention_probs_dropout_prob=cfg.get("attention_probs_dropout_prob", 0.1))
    acfg = types.SimpleNamespace(hidden_size=H, num_attention_heads=nh, position_embeddings_type="relative_key",
                                 left_max_position_embeddings=P - 1, right_max_position_embeddings=P - 1,
                                 attention_dropout=bcfg.attention_probs_dropout_prob)

    class Layer(nn.Module):
        def __init__(self):
            super().__init__()
            self.att = hf_w2v.Wav2Vec2BertSelfAttention(acfg)
            self.self_out = hf_bert.BertSelfOutput(bcfg)
            self.inter = hf_bert.BertIntermediate(bcfg)
            self.out = hf_bert.BertOutput(bcfg)

        def forward(self, h, ext_mask):
            ctx, _ = self.att(h, attention_mask=ext_mask)
            a = self.self_out(ctx, h)
            return self.out(self.inter(a), a)

    class ReferenceAssembled(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = bcfg
            self.n_inputs = n_in
            self.inputs_to_hidden_dim = nn.Linear(n_in, H)
            self.embeddings = rm.BertEmbeddings(bcfg)            # the reference's class
            self.time_embed = rm.GaussianFourierProjection(H)    # the reference's class
            self.layers = nn.ModuleList(Layer() for _ in range(L))
            self.token_decoder = rm.AnglesPredictor(H, n_in)     # the reference's class

        def forward(self, inputs, timestep, attention_mask, position_ids=None):
            # modelling.py:427-484
            input_shape = inputs.size()
            batch_size, seq_length, *_ = input_shape
            assert attention_mask.dim() == 2
            ext = attention_mask[:, None, None, :].type_as(attention_mask)
            ext = (1.0 - ext) * -10000.0
            if position_ids is None:
                position_ids = torch.arange(seq_length).expand(batch_size, -1)
            h = self.embeddings(self.inputs_to_hidden_dim(inputs), position_ids=position_ids)
            h = h + self.time_embed(timestep.squeeze(dim=-1)).unsqueeze(1)
            for layer in self.layers:
                h = layer(h, ext)
            return self.token_decoder(h)

    m = ReferenceAssembled()
    with torch.no_grad():
        own = m.state_dict()
        direct = ["inputs_to_hidden_dim.weight", "inputs_to_hidden_dim.bias", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias",
                  "time_embed.W", "token_decoder.dense1.weight", "token_decoder.dense1.bias", "token_decoder.layer_norm.weight",
                  "token_decoder.layer_norm.bias", "token_decoder.dense2.weight", "token_decoder.dense2.bias"]
        for k in direct:
            own[k].copy_(state_dict[k])
        for l in range(L):
            p, q = f"encoder.layer.{l}.", f"layers.{l}."
            pairs = {"att.linear_q": "attention.self.query", "att.linear_k": "attention.self.key", "att.linear_v": "attention.self.value",
                     "att.linear_out": "attention.output.dense", "inter.dense": "intermediate.dense", "out.dense": "output.dense"}
            for mine, theirs in pairs.items():
                own[q + mine + ".weight"].copy_(state_dict[p + theirs + ".weight"])
                own[q + mine + ".bias"].copy_(state_dict[p + theirs + ".bias"])
            own[q + "att.distance_embedding.weight"].copy_(state_dict[p + "attention.self.distance_embedding.weight"].flip(0))
            own[q + "self_out.dense.weight"].copy_(torch.eye(H))  # the output projection already ran inside `att`
            own[q + "self_out.dense.bias"].zero_()
            for mine, theirs in {"self_out.LayerNorm": "attention.output.LayerNorm", "out.LayerNorm": "output.LayerNorm"}.items():
                own[q + mine + ".weight"].copy_(state_dict[p + theirs + ".weight"])
                own[q + mine + ".bias"].copy_(state_dict[p + theirs + ".bias"])
    return m.eval()
