"""Stage-wise error attribution of the CUDA path against the CPU oracle (run on the GPU box: python tests/diag_parity.py).
Test infrastructure: lives under tests/ because it imports oracle/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_grad_enabled(False)

import npvp_b200  # noqa: E402
from npvp_b200.config import RENORM  # noqa: E402
from npvp_b200.pipeline import build_from_config  # noqa: E402
from oracle import npvp_oracle as O  # noqa: E402
from util_init import reset_shared_norm, seeded_rand, seeded_randn, stress_init_  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def main():
    cases = [("KITTI_VFP_NPVP-S", 2, True), ("SMMNIST_VFP_NPVP-D", 2, True), ("Cityscapes_VFP_NPVP-S", 1, False),
             ("BAIR_VFP_NPVP-S", 2, True), ("BAIR_VFP_NPVP-S", 2, False), ("KTH_Unified_NPVP-S", 2, True)]
    for name, N, stress in cases:
        reset_shared_norm(npvp_b200.Predictor)
        model = build_from_config(name, device="cpu", seed=0)
        if stress:
            stress_init_(model.VPTR_Enc, 1); stress_init_(model.VPTR_Dec, 2); stress_init_(model.predictor, 3)
        cfg = model.cfg
        To, hw, ch = cfg.Dataset.num_past_frames, cfg.Dataset.img_size, cfg.Dataset.img_channels
        x = seeded_rand((N, To, ch, hw, hw), 1234)
        if cfg.AE.out_layer == "Tanh":
            x = x * 2 - 1
        stoch = cfg.Predictor.stochastic
        eps = seeded_randn((N, 512, 8, 8), 4321)
        esd, psd, dsd = model.VPTR_Enc.state_dict(), model.predictor.state_dict(), model.VPTR_Dec.state_dict()
        nd, nr, outl = cfg.AE.n_downsampling, cfg.AE.num_res_blocks, cfg.AE.out_layer
        f_ref = O.resnet_encoder(esd, x, nd, nr)
        oc, pc = model.predictor.observed_coor, model.predictor.predict_coor      # (buffers are absent from the state_dict with rand_context)
        p_ref = O.predictor_forward(psd, f_ref, oc, pc, stoch, eps if stoch else None)
        o_ref = O.resnet_decoder(dsd, p_ref, nd, outl)
        model = model.cuda()
        model.predictor.injected_eps = eps.cuda() if stoch else None
        f = model.VPTR_Enc(x.cuda())
        p_from_ref = model.predictor(f_ref.cuda())
        p = model.predictor(f)
        o_from_ref = model.VPTR_Dec(p_ref.cuda())
        o = model.VPTR_Dec(p)
        std = max(RENORM[cfg.Dataset.name][1]) if outl == "Tanh" else 1.0
        print(f"{name} stress={stress} std={std:.2f} | enc rel {rel(f.cpu(), f_ref):.2e} | pred(ref feats) rel {rel(p_from_ref.cpu(), p_ref):.2e} "
              f"| pred(chain) rel {rel(p.cpu(), p_ref):.2e} | dec(ref feats) abs {float((o_from_ref.cpu() - o_ref).abs().max()):.2e} "
              f"| e2e abs {float((o.cpu() - o_ref).abs().max()):.2e} -> pixel {std * float((o.cpu() - o_ref).abs().max()):.2e} "
              f"| out range [{float(o_ref.min()):.2f},{float(o_ref.max()):.2f}] feat max {float(p_ref.abs().max()):.1f}", flush=True)
        model.cpu()


if __name__ == "__main__":
    main()
