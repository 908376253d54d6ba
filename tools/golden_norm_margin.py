"""How far the gradient norms of the drop-in are from the reference's golden norms, per sweep mode (SPAIR_SWEEP_TC = 0 / bwd / 1):
the margin against the 1e-4 bound of tests/helpers.check_model_against_golden.  python tools/golden_norm_margin.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers  # noqa: E402


def main():
    for name, step in (("tiny", 1), ("tiny", 1001), ("A", 1), ("A", 1001)):
        g = helpers.load_golden("model_%s_step%d.npz" % (name, step))
        for mode in ("0", "bwd", "1"):
            os.environ["SPAIR_SWEEP_TC"] = mode
            net = helpers.build_model(name, "cuda")
            x = torch.from_numpy(g["x"]).cuda()
            net.set_noise(*(torch.from_numpy(g[k]) for k in ("eps_where", "eps_attr", "eps_depth", "u_pres")))
            loss = net(x, int(g["step"]))[0]
            loss.backward()
            worst, where = 0.0, ""
            for k, p in net.named_parameters():
                if p.grad is None:
                    continue
                ref = float(g["gstat/" + k][1])
                dev = abs(p.grad.double().norm().item() - ref) / max(ref, 1e-30)
                if dev > worst:
                    worst, where = dev, k
            print("%-5s step %-5d SPAIR_SWEEP_TC=%-3s  worst |norm - ref| / ref = %.3e  (%s)   loss %.6f vs %.6f"
                  % (name, step, mode, worst, where, float(loss), float(g["loss"])))


if __name__ == "__main__":
    main()
