"""GPU diagnostic for the decoder path: layer-by-layer forward and backward comparison with the oracle."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tests.helpers import (DEC_BACKWARD_STAGE_NAMES, cosine, fetch_decoder_saved, oracle_decoder_trace, rel_err)  # noqa


def run_decoder(dev, res, batch):
    from oracle.decoder_oracle import make_vae_oracle
    from oracle.encoder_oracle import perturb_affine_params
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = make_vae_oracle(0)
    perturb_affine_params(model, 1234)
    g = torch.Generator().manual_seed(200 + res)
    h = res // 8
    z = torch.randn((batch, 4, h, h), generator=g)
    dimg = torch.randn((batch, 3, res, res), generator=g) * 1e-2
    md = model.to(dev)
    acts, grads, img_ref, dz_ref = oracle_decoder_trace(md, z.to(dev), dimg.to(dev))
    model.to("cpu")
    vae = AutoencoderKL(device=str(dev)).load_state_dict(model.state_dict())
    img, saved = vae._decode_raw(z.to(dev), keep=True)
    torch.cuda.synchronize()
    ok = True

    def report(name, a, b, tol):
        nonlocal ok
        e, c = rel_err(a, b), cosine(a, b)
        good = e < tol and not torch.isnan(a).any()
        ok &= bool(good)
        print(f"[dec] {name:12s} rel_err={e:.3e} cos={c:.6f} {'OK' if good else 'FAIL'}", flush=True)

    report("conv_in", fetch_decoder_saved(vae, saved, "conv_in"), acts["conv_in"], 2e-2)
    for i in range(14):
        report(f"res{i}_out", fetch_decoder_saved(vae, saved, "resnet_out", i), acts[f"res{i}_out"], 4e-2)
        if i == 0:
            report("attn_out", fetch_decoder_saved(vae, saved, "attn_out"), acts["attn_out"], 4e-2)
        if i in (4, 7, 10):
            k = (i - 4) // 3
            report(f"up{k}_out", fetch_decoder_saved(vae, saved, "up_out", k), acts[f"up{k}_out"], 4e-2)
    report("image", img, img_ref, 4e-2)

    slot = batch * res * res * 256 * 2
    dump = torch.zeros(len(DEC_BACKWARD_STAGE_NAMES) * slot, dtype=torch.uint8, device=dev)
    vae._lib.tml_debug_set_grad_dump(dump.data_ptr(), slot, len(DEC_BACKWARD_STAGE_NAMES))
    dz = vae._decode_backward_raw(dimg.to(dev), saved, tuple(z.shape))
    torch.cuda.synchronize()
    vae._lib.tml_debug_set_grad_dump(None, 0, 0)
    for k, name in enumerate(DEC_BACKWARD_STAGE_NAMES):
        ref = grads[name]
        Bn, Cc, Hh, Ww = ref.shape
        n = Bn * Cc * Hh * Ww
        t = dump[k * slot: k * slot + 2 * n].view(torch.bfloat16).view(Bn, Hh, Ww, Cc).float().permute(0, 3, 1, 2)
        report("d_" + name, t, ref, 8e-2)
    report("dz", dz, dz_ref, 8e-2)
    print(f"[dec] dz cosine = {cosine(dz, dz_ref):.6f}")
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    ok = run_decoder(torch.device("cuda:0"), a.res, a.batch)
    print("ALL OK" if ok else "SOME CHECKS FAILED")
    sys.exit(0 if ok else 1)
