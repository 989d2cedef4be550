"""tests/golden/psp_encoder.npz: the UNMODIFIED reference GradualStyleEncoder (IR-SE50) run on CPU.

    python -m oracle.make_golden_psp          (build container only: needs /root/reference)
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import psp_state, ref_shim  # noqa: E402


def main():
    ref = ref_shim.load_reference_psp_encoders()

    class Opts:
        input_nc = 3

    enc = psp_state.randomize(psp_state.build(ref.GradualStyleEncoder, opts=Opts()))
    x = psp_state.image()
    with torch.no_grad(), ref_shim.cpu_mode():
        z, p = enc(x)
    out = os.path.join(ROOT, "tests", "golden", "psp_encoder.npz")
    np.savez_compressed(out, z=z.numpy(), p=p.numpy(), checksum=np.float64(psp_state.checksum(enc)),
                        keys=np.array(sorted(enc.state_dict().keys())))
    print("wrote", out, "z", tuple(z.shape), float(z.abs().mean()), "p", tuple(p.shape), float(p.abs().mean()))


if __name__ == "__main__":
    main()
