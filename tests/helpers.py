"""Helpers shared by the test modules."""
import torch


def digest_err(t, dg):
    t = t.detach().contiguous().reshape(-1).double().cpu()
    idx = torch.tensor(dg['idx'])
    val = torch.tensor(dg['val'], dtype=torch.float64)
    scale = max(dg['norm'] / max(t.numel(), 1) ** 0.5, 1e-30)
    e_samples = float((t[idx] - val).abs().max() / scale)
    e_norm = abs(float(t.norm()) - dg['norm']) / max(dg['norm'], 1e-30)
    e_sum = abs(float(t.sum()) - dg['sum']) / max(dg['abssum'], 1e-30)
    return max(e_samples, e_norm, e_sum)


# ---- parity report: every GPU parity test records its measured errors; conftest dumps them ----
REPORT = {}


def record(test, case, **vals):
    REPORT.setdefault(test, {})[case] = {k: (float(v) if not isinstance(v, (int, str)) else v) for k, v in vals.items()}
