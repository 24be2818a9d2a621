"""ncu target: ONE eager (ungraphed) SDXL UNet forward at BASELINE config 2 between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_forward.py

Weights stream from HBM exactly as in a real step (two warm-up forwards run first, 5.1 GB of weights each, so
nothing but the activations is L2-resident)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stabletriton_b200 as st  # noqa: E402
from stabletriton_b200 import UNetConfig, synth  # noqa: E402


def main():
    cfg = UNetConfig.tiny() if "--tiny" in sys.argv else UNetConfig.sdxl()
    latent = 32 if "--tiny" in sys.argv else 128
    model = synth.build_unet(cfg, seed=7)
    compiled = st.compile(model, cuda_graph=False)
    inp = synth.synth_inputs(2, latent, cfg, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(2):
            compiled(**inp)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        compiled(**inp)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("one forward done")


if __name__ == "__main__":
    main()
