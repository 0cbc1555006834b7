"""Diagnostic: clock64 deltas of one sample of the fused forward kernel (STC_FUSED_TRACE=1)."""
import os, sys, torch
os.environ["STC_FUSED_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stc_gnn_b200 as S
from stc_gnn_b200.synth import sf_supports
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Din = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Gs, Gc = sf_supports()
cell = S.STC_Cell(100, 5, 2, 2, Din, 16).to(dev)
X = torch.randn(B, 100, 5, Din, device=dev); H = torch.randn(B, 100, 5, 16, device=dev)
with torch.no_grad():
    for _ in range(2):
        out = cell(Gs=Gs.to(dev), Gc=Gc.to(dev), Xt=X, Ht_1=H)
torch.cuda.synchronize()
