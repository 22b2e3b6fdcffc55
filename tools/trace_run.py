"""One traced forward of an ACE2-size network (development): writes the tcgen05 kernel's per-tile clock samples to
$ACE_B200_TRACE_FILE (default gpurun_out/trace.bin); summarise with tools/trace_report.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ACE_B200_TRACE_FILE", "gpurun_out/trace.bin")
import torch
from ace_b200 import _lib
from gemm_probe import build, IMG

if __name__ == "__main__":
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    opts = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    if os.path.exists(os.environ["ACE_B200_TRACE_FILE"]):
        os.remove(os.environ["ACE_B200_TRACE_FILE"])
    net = build(layers=layers)
    x = torch.randn(1, 44, *IMG, device="cuda")
    for k, v in opts.items():
        _lib.set_option(k, v)
    with torch.no_grad():
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        _lib.set_option("trace", 1)
        net(x)
        torch.cuda.synchronize()
        _lib.set_option("trace", 0)
