"""Generate the golden fixtures under tests/golden/ (run in the BUILD container, where
/root/reference is mounted; the GPU box never runs this).

The reference's ``main.py`` cannot be imported (``diffusers``, ``wandb`` login, ... are missing),
so the functions on the hot path are pulled out of its source with ``ast`` and executed as is:

* ``Trainer.perturbation_step``                      (main.py:248-276)   -> pgd_linf_*.npz, pgd_l2_*.npz
* the universal-delta update statements               (old/train_noise.py:172-185) -> universal_update.npz
* ``losses/losses.py`` (importable: depends on torch only)                -> losses.npz

No reference source text is written to the repo: only the numeric inputs/outputs.
Encoder fixtures (``encoder_*.npz``) come from ``oracle/encoder_oracle.py`` itself (regression
fixtures: the encoder lives in the absent ``diffusers`` package, parity there is unpinned).

Usage:  python -m oracle.gen_golden [--ref /root/reference] [--out tests/golden]
"""
from __future__ import annotations

import argparse
import ast
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch


def _extract_method(path: Path, cls: str, name: str):
    """Compile one method of one class of a reference file into a standalone function."""
    tree = ast.parse(path.read_text())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    fn.decorator_list = []
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = {"torch": torch, "np": np}
                    exec(compile(mod, str(path), "exec"), ns)
                    return ns[name]
    raise LookupError(f"{cls}.{name} not found in {path}")


def _extract_statements(path: Path, func: str, first_line: int, last_line: int):
    """Compile the statements of ``func`` whose lines fall in [first_line, last_line]."""
    tree = ast.parse(path.read_text())
    out = []

    def visit(stmts):
        for s in stmts:
            if first_line <= s.lineno and getattr(s, "end_lineno", s.lineno) <= last_line:
                out.append(s)
            else:
                for fld in ("body", "orelse"):
                    sub = getattr(s, fld, None)
                    if isinstance(sub, list):
                        visit(sub)

    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == func:
            visit(node.body)
    if not out:
        raise LookupError(f"no statements in {path}:{first_line}-{last_line}")
    mod = ast.Module(body=out, type_ignores=[])
    ast.fix_missing_locations(mod)
    return compile(mod, str(path), "exec")


def special_linf_inputs(gen: torch.Generator, shape, eps, step, lo, hi):
    """Random tensors salted with the edge cases SURVEY section 4 lists."""
    x = torch.rand(shape, generator=gen) * (hi - lo) + lo
    x_adv = (x + (torch.rand(shape, generator=gen) * 2 - 1) * eps).clamp(lo, hi)
    grad = torch.randn(shape, generator=gen) * 1e-3
    flat_g, flat_x, flat_a = grad.view(-1), x.view(-1), x_adv.view(-1)
    specials = [0.0, -0.0, 1e-30, -1e-30, float("nan"), float("inf"), -float("inf"), 1e-45, -1e-45]
    for i, v in enumerate(specials):
        flat_g[i] = v
    n = flat_g.numel()
    # exactly on the eps boundary / clamp bounds, about to cross them
    flat_a[20] = flat_x[20] + eps
    flat_g[20] = -1.0
    flat_a[21] = flat_x[21] - eps
    flat_g[21] = 1.0
    flat_x[22] = hi
    flat_a[22] = hi
    flat_g[22] = -1.0
    flat_x[23] = lo
    flat_a[23] = lo
    flat_g[23] = 1.0
    flat_x[24] = hi - eps / 2
    flat_a[24] = hi
    flat_g[24] = -1.0
    flat_a[n - 1] = float("nan")
    return x_adv, grad, x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=str(Path(__file__).resolve().parents[1] / "tests" / "golden"))
    args = ap.parse_args()
    ref, out = Path(args.ref), Path(args.out)
    out.mkdir(parents=True, exist_ok=True)

    # ---------------------------------------------------------------- perturbation_step (main.py)
    step_fn = _extract_method(ref / "main.py", "Trainer", "perturbation_step")

    def run_step(norm_type, eps, step, lo, hi, x_adv, grad, x, mask):
        cfg = types.SimpleNamespace(norm_type=norm_type, eps=eps, step_size=step, min_value=lo, max_value=hi)
        me = types.SimpleNamespace(cfg=cfg)
        return step_fn(me, X_adv=x_adv.clone(), grad=grad.clone(), X=x.clone(), X_mask=mask)

    gen = torch.Generator().manual_seed(0)
    linf_cases = {
        "refdefault": (0.1, 0.006, -1, 1),                 # configs.py:157-158, :127-129
        "northstar": (32 / 255, 4 / 255, -1, 1),           # eps 16/255, step 2/255 on a [0,1] scale
        "unit": (16 / 255, 2 / 255, 0, 1),
        "playground": (16, 1, -1, 1),                      # old/yuval_playground.py:364-365
    }
    for name, (eps, step, lo, hi) in linf_cases.items():
        x_adv, grad, x = special_linf_inputs(gen, (2, 3, 16, 24), eps, step, lo, hi)
        y = run_step("linf", eps, step, lo, hi, x_adv, grad, x, None)
        np.savez_compressed(out / f"pgd_linf_{name}.npz", x_adv=x_adv.numpy(), grad=grad.numpy(), x=x.numpy(),
                            params=np.array([eps, step, lo, hi], dtype=np.float64), out=y.numpy())

    l2_cases = {
        "refdefault": (32.0, 7.5, -1, 1, False),           # configs.py:153-154
        "small_eps": (2.0, 1.5, -1, 1, False),             # projection active for every image
        "masked": (4.0, 2.0, -1, 1, True),
    }
    for name, (eps, step, lo, hi, use_mask) in l2_cases.items():
        shape = (3, 3, 16, 24)
        x = torch.rand(shape, generator=gen) * 2 - 1
        x_adv = (x + 0.05 * torch.randn(shape, generator=gen)).clamp(-1, 1)
        x_adv[1] = x[1]                                    # zero perturbation row
        grad = torch.randn(shape, generator=gen) * 1e-3
        grad[2] = 0.0                                      # zero gradient row: g/(0+1e-10)
        mask = (torch.rand((3, 1, 16, 24), generator=gen) > 0.5).float() if use_mask else None
        y = run_step("l2", eps, step, lo, hi, x_adv, grad, x, mask)
        np.savez_compressed(out / f"pgd_l2_{name}.npz", x_adv=x_adv.numpy(), grad=grad.numpy(), x=x.numpy(),
                            mask=(mask.numpy() if mask is not None else np.zeros(0, np.float32)),
                            params=np.array([eps, step, lo, hi], dtype=np.float64), out=y.numpy())

    # ---------------------------------------------------------------- universal update (train_noise.py)
    code = _extract_statements(ref / "old" / "train_noise.py", "main", 172, 185)
    for name, (eps, step, shape) in {"ref": (16.0, 1.0, (1, 3, 16, 24)), "tight": (0.05, 0.5, (1, 3, 16, 24))}.items():
        src = torch.rand(shape, generator=gen) * 2 - 1
        delta0 = torch.randn(shape, generator=gen) * 0.02
        grad = torch.randn(shape, generator=gen) * 1e-3
        ns = {"torch": torch, "grad": grad.clone(), "source_image": src.clone(),
              "perturbation": delta0.clone().requires_grad_(True),
              "cfg": types.SimpleNamespace(step_size=step, eps=eps, apply_image_pertubation=True)}
        exec(code, ns)
        np.savez_compressed(out / f"universal_update_{name}.npz", delta=delta0.numpy(), grad=grad.numpy(),
                            source=src.numpy(), params=np.array([eps, step], dtype=np.float64),
                            out=ns["perturbation"].detach().numpy())

    # ---------------------------------------------------------------- losses (losses/losses.py)
    spec = importlib.util.spec_from_file_location("_ref_losses", ref / "losses" / "losses.py")
    ref_losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_losses)
    a = torch.randn((2, 4, 8, 8), generator=gen)
    b = torch.randn((2, 4, 8, 8), generator=gen)
    np.savez_compressed(
        out / "losses.npz", a=a.numpy(), b=b.numpy(),
        perturbation_loss=ref_losses.perturbation_loss(a, b).numpy(),
        l2_distance=ref_losses.LpDistance(2)(a, b).numpy(),
        linf_distance=ref_losses.LpDistance(float("inf"))(a, b).numpy(),
        l2_regularization=ref_losses.LpRegularization(2)([a, b]).numpy(),
        cosine=ref_losses.CosineSimilarity()(a, b).numpy())

    # ---------------------------------------------------------------- encoder regression fixtures
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle.encoder_oracle import make_oracle, perturb_affine_params, encoder_attack_grad
    torch.set_num_threads(8)
    model = make_oracle(0)
    perturb_affine_params(model, 1234)
    assert sum(p.numel() for p in model.parameters()) == 34_163_664
    for res in (64, 128):
        g = torch.Generator().manual_seed(100 + res)
        x = torch.rand((2, 3, res, res), generator=g) * 2 - 1
        tgt = torch.randn((2, 4, res // 8, res // 8), generator=g)
        noise = torch.randn((2, 4, res // 8, res // 8), generator=g)
        with torch.no_grad():
            moments = model.moments(x)
        out_d = {"x": x.numpy(), "target": tgt.numpy(), "noise": noise.numpy(), "moments": moments.numpy()}
        for kind in (0, 1):
            gr, ls, z = encoder_attack_grad(model, x, tgt, noise, kind)
            out_d[f"grad_kind{kind}"] = gr.numpy()
            out_d[f"loss_kind{kind}"] = ls.numpy()
            out_d["z"] = z.numpy()
        np.savez_compressed(out / f"encoder_{res}.npz", **out_d)
    # ---------------------------------------------------------------- decoder regression fixture
    from oracle.decoder_oracle import make_vae_oracle, autoencoder_attack_grad
    vae = make_vae_oracle(0)
    perturb_affine_params(vae, 1234)
    assert sum(p.numel() for p in vae.parameters()) == 83_653_863
    g = torch.Generator().manual_seed(300)
    x = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    tgt_img = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    noise = torch.randn((1, 4, 8, 8), generator=g)
    z = torch.randn((1, 4, 8, 8), generator=g)
    with torch.no_grad():
        img = vae.decode(z)
    gr, ls, outimg = autoencoder_attack_grad(vae, x, tgt_img, x, noise, 1.0, 1.0)
    np.savez_compressed(out / "decoder_64.npz", z=z.numpy(), image=img.numpy(), x=x.numpy(), target_image=tgt_img.numpy(),
                        noise=noise.numpy(), grad=gr.numpy(), loss=ls.numpy(), output_image=outimg.numpy())
    print("golden fixtures written to", out)


if __name__ == "__main__":
    main()
