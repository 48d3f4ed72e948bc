"""debug: one golden model case, every quantized layer teacher-forced; prints the deviation of the engine AND of the fp32
CPU oracle from an fp64 replay of the oracle layer (whose error is it?).  Development aid (uses the oracle)."""
import copy, sys
import torch
from tests.golden.cases import MODEL_CASES
from tests.oracle_util import load_golden, rel_err
from tests.test_gpu_parity import _zoo_model, _prepare_engine, QUANT_TYPES, DEV
from tests.test_oracle_golden import prepare_oracle

name = sys.argv[1]
case = next(c for c in MODEL_CASES if c["name"] == name)
gold = load_golden("model", name)
init = {k[5:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("init.")}
om = _zoo_model(case); om.load_state_dict(init); om = prepare_oracle(om, case); om.train()
em = _zoo_model(case); em.load_state_dict(init); em = _prepare_engine(em, case).to(DEV); em.train()
pristine = copy.deepcopy(om)
x, t = torch.from_numpy(gold["s0.x"]), torch.from_numpy(gold["s0.t"])
names = [n for n, m in em.named_modules() if type(m).__name__ in QUANT_TYPES and not n.endswith("activation_quantizer")]
cap = {}
def fh(n):
    def h(mod, inp, out):
        rec = {"x": [t.detach().clone() for t in inp], "y": out.detach().clone()}
        cap[n] = rec
        out.register_hook(lambda g: rec.__setitem__("go", g.detach().clone()))
    return h
omods = dict(om.named_modules())
hs = [omods[n].register_forward_hook(fh(n)) for n in names]
torch.nn.functional.cross_entropy(om(x), t).backward()
for h in hs: h.remove()
emods, pmods = dict(em.named_modules()), dict(pristine.named_modules())
for n in names:
    e, c = emods[n], cap[n]
    o32, o64 = copy.deepcopy(pmods[n]), copy.deepcopy(pmods[n]).double()
    x32 = [t.clone().requires_grad_(True) for t in c["x"]]
    x64 = [t.double().clone().requires_grad_(True) for t in c["x"]]
    o32(*x32).backward(c["go"]); o64(*x64).backward(c["go"].double())
    xe = [t.to(DEV).requires_grad_(True) for t in c["x"]]
    e.zero_grad(); e(*xe).backward(c["go"].to(DEV))
    g64 = {k: p.grad for k, p in o64.named_parameters()}
    g32 = {k: p.grad for k, p in o32.named_parameters()}
    for k, p in e.named_parameters():
        if g64.get(k) is None: continue
        print(f"{n}.{k}: engine-vs-f64 {rel_err(p.grad, g64[k]):.2e}  oracle32-vs-f64 {rel_err(g32[k], g64[k]):.2e}  engine-vs-oracle32 {rel_err(p.grad, g32[k]):.2e}")
    for i in range(len(xe)):
        if x64[i].grad is not None and xe[i].grad is not None:
            print(f"{n}.dx{i}: engine-vs-f64 {rel_err(xe[i].grad, x64[i].grad):.2e}  oracle32-vs-f64 {rel_err(x32[i].grad, x64[i].grad):.2e}")
