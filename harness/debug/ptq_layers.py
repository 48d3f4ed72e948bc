"""debug: eval-mode (PTQ) engine modules fed the oracle's inputs, layer by layer; prints the deviation of every quantized
module's output and of every calibrated buffer.  Uses the oracle: development aid only, never imported by the product."""
import copy, sys
import torch
from tests.test_gpu_inference import _prepared, DEV
from tests.oracle_util import rel_err
from tests.test_gpu_parity import QUANT_TYPES

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 224
eng, ora, calib, x = _prepared((64, 128, 256, 512), hw)
with torch.no_grad():
    eng.train(); eng(calib.to(DEV)); eng.eval()
    ora.train(); ora(calib); ora.eval()
    so, se = ora.state_dict(), eng.state_dict()
    for k in so:
        if so[k].dtype.is_floating_point and k in se:
            e = rel_err(se[k], so[k])
            if e > 1e-6:
                print(f"buffer {k}: {e:.2e}")
    names = [n for n, m in eng.named_modules() if type(m).__name__ in QUANT_TYPES and not n.endswith("activation_quantizer")]
    cap = {}
    def hook(n):
        def h(mod, inp, out):
            cap[n] = ([t.detach().clone() for t in inp], out.detach().clone())
        return h
    om = dict(ora.named_modules())
    hs = [om[n].register_forward_hook(hook(n)) for n in names]
    yo = ora(x)
    for h in hs: h.remove()
    em = dict(eng.named_modules())
    for n in names:
        xin, yref = cap[n]
        ye = em[n](*[t.to(DEV) for t in xin]).cpu()
        d = (ye - yref).abs()
        print(f"{n:40s} {type(em[n]).__name__:22s} rel {rel_err(ye, yref):.2e}  frac>1e-4: {(d > 1e-4 * yref.abs().max()).float().mean().item():.2e}")
    ye = eng(x.to(DEV)).cpu()
    print("logits plain", rel_err(ye, yo))
    from micronet_b200 import iao
    iao.freeze_inference(eng)
    print("logits frozen", rel_err(eng(x.to(DEV)).cpu(), yo))
