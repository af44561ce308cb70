"""Shared helpers of the -m gpu tests (the parity tests proper; they call through the C ABI)."""
import torch

from foldingdiff_b200 import modelling, synthetic

GEMMS = ["fp32", "tc3x"]
# max-abs tolerance on eps_hat against the fp32 CPU oracle, per GEMM arithmetic
FWD_TOL = {"fp32": 1e-5, "tc3x": 1e-5}  # measured (B200): 1.1e-6 fp32, 1.5e-6 tc3x at production shape


def mini_model(mini_dir, gemm):
    return modelling.BertForDiffusionBase.from_dir(mini_dir).to("cuda:0").set_gemm(gemm)


_PROD_SD = None


def prod_state_dict():
    global _PROD_SD
    if _PROD_SD is None:
        _PROD_SD = synthetic.synthetic_state_dict(synthetic.PRODUCTION, seed=0)
    return _PROD_SD


def prod_model(gemm):
    cfg = modelling.BertConfig(**synthetic.PRODUCTION)
    m = modelling.BertForDiffusionBase(cfg, ft_is_angular=[True] * 6)
    m.load_state_dict(prod_state_dict())
    return m.to("cuda:0").set_gemm(gemm)


def prefix_mask(lengths, n):
    mask = torch.zeros(len(lengths), n)
    for i, l in enumerate(lengths):
        mask[i, :l] = 1.0
    return mask
