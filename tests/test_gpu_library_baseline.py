"""Informational, off by default: the "library" competitor SURVEY 8(d) / BASELINE.md 3 ask to time beside the kernels --
the same encoder attack through PyTorch on the B200 (bf16 autocast, channels_last, cuDNN convolutions, SDPA attention,
torch.autograd for the input gradient).  Not a parity test; run it with

    TML_LIBRARY_BASELINE=1 python -m pytest tests/test_gpu_library_baseline.py -m gpu -s -q

It prints one JSON line (and writes it to gpurun_out/library_baseline.json when that directory exists)."""
import json
import os
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _sdpa_forward(self, x):
    B, C, H, W = x.shape
    r = x
    t = self.group_norm(x.reshape(B, C, H * W)).transpose(1, 2)
    q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
    a = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)   # one head, d = C
    o = self.to_out[0](a)
    return o.transpose(1, 2).reshape(B, C, H, W) + r


@pytest.mark.skipif(os.environ.get("TML_LIBRARY_BASELINE") != "1", reason="informational timing; set TML_LIBRARY_BASELINE=1")
def test_library_baseline_report():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import encoder_oracle as eo
    from tml_image_editing_defense_b200 import ops
    dev = torch.device("cuda:0")
    res, B, steps, warm = 512, int(os.environ.get("TML_LIBRARY_BATCH", "16")), 5, 2
    model = eo.make_oracle(0).to(dev).to(memory_format=torch.channels_last)
    model.requires_grad_(False)
    original_forward = eo.Attention.forward
    eo.Attention.forward = _sdpa_forward          # (restored below: other tests use the oracle's own attention)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand((B, 3, res, res), generator=g) * 2 - 1).to(dev)
    tgt = torch.randn((B, 4, res // 8, res // 8), generator=g).to(dev)
    noise = torch.randn((B, 4, res // 8, res // 8), generator=g).to(dev)
    x_adv = x.clone()

    def step():
        xx = x_adv.detach().clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            m = model.moments(xx).float()
        mean, logvar = m.chunk(2, 1)
        z = mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise
        loss = (z - tgt).flatten(1).norm(dim=1).sum()
        (grad,) = torch.autograd.grad(loss, [xx])
        ops.pgd_step_linf_(x_adv, grad.float().contiguous(), x, 32 / 255, 4 / 255, -1.0, 1.0)

    try:
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    finally:
        eo.Attention.forward = original_forward
    ms = e0.elapsed_time(e1) / steps
    line = {"impl": "pytorch library path on the same B200 (bf16 autocast, channels_last, cuDNN, SDPA, autograd)",
            "metric": "image-PGD-iters/sec", "value": B / (ms * 1e-3), "ms_per_step": ms, "batch": B, "resolution": res,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "torch": torch.__version__,
            "cudnn": torch.backends.cudnn.version()}
    print(json.dumps(line))
    out = ROOT / "gpurun_out"
    if out.exists():
        (out / "library_baseline.json").write_text(json.dumps(line) + "\n")
