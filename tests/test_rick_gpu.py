"""Bit-exact parity of the device Fisher -> FIM -> percentile -> mask step (rick_b200/rick.py over the C ABI) against
the golden index sets produced by executing the reference's own code, and against the NumPy oracle."""
import numpy as np
import pytest
import torch

from oracle import rick_oracle as ro
from oracle import synth

pytestmark = pytest.mark.gpu


def _unpack(bits, n):
    return np.unpackbits(bits)[:n].astype(bool)


def _to_cuda(f):
    return {k: torch.from_numpy(v).cuda() for k, v in f.items()}


def test_masks_bit_exact_vs_reference_golden(golden):
    from rick_b200 import rick
    gold = golden("rick_masks_golden.npz")
    masks_g = masks_d = None
    for rnd, (sg, sd) in enumerate([(101, 102), (103, 104)]):
        fg, fd = synth.fisher_g(sg), synth.fisher_d(sd)
        cg, cd = _to_cuda(fg), _to_cuda(fd)
        if masks_g is None:
            masks_g = rick.FilterMasks(rick.generator_layers(cg), "cuda")
            masks_d = rick.FilterMasks(rick.discriminator_layers(cd), "cuda")
        masks_g.update(cg, 40.0, 0.1)
        masks_d.update(cd, 40.0, 0.1)
        want = gold[f"r{rnd}/lines"]
        got = np.concatenate([masks_g.lines["conv"].cpu().numpy(), masks_g.lines["fc"].cpu().numpy(),
                              masks_d.lines["d"].cpu().numpy()])
        assert np.array_equal(got, want), (got, want)           # float64 thresholds, bit for bit
        for masks, suffix, f in ((masks_g, "g", fg), (masks_d, "d", fd)):
            fr, ft, pr, zero = masks.index_sets()
            for tag, sets in ((f"freeze_{suffix}", fr), (f"ft_{suffix}", ft), (f"prune_{suffix}", pr),
                              (f"zero_{suffix}", zero)):
                keys = [k for k in gold.files if k.startswith(f"r{rnd}/{tag}/")]
                assert len(keys) == len(sets) > 0
                for k in keys:
                    name = k.split("/", 2)[2]
                    n = f[name].shape[1] if f[name].ndim == 5 else f[name].shape[0]
                    m = np.zeros(n, bool)
                    m[sets[name]] = True
                    assert np.array_equal(m, _unpack(gold[k], n)), k


@pytest.mark.parametrize("rows,length", [(512, 4608), (256, 1152), (512, 512), (128, 27), (64, 7), (33, 129),
                                         (17, 1000), (3, 20000), (5, 1), (130, 8), (64, 128), (64, 136)])
def test_filter_fim_equals_numpy_mean_bitwise(rows, length):
    from rick_b200 import _lib
    rng = np.random.default_rng(rows * 31 + length)
    a = (rng.standard_normal((rows, length), dtype=np.float32) ** 2 * 1e-6).astype(np.float32)
    b = (rng.standard_normal(rows, dtype=np.float32) ** 2 * 1e-6).astype(np.float32)
    ca, cb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty(rows, device="cuda")
    lib = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.rick_filter_fim(out.data_ptr(), ca.data_ptr(), None, rows, length, s), "fim")
    assert np.array_equal(out.cpu().numpy(), a.mean(axis=1))
    _lib.check(lib.rick_filter_fim(out.data_ptr(), ca.data_ptr(), cb.data_ptr(), rows, length, s), "fim")
    assert np.array_equal(out.cpu().numpy(), (a.mean(axis=1) + b) / 2)


@pytest.mark.parametrize("n", [1, 2, 3, 10, 257, 4864, 8064, 100000])
def test_percentile_equals_numpy_bitwise(n):
    import ctypes
    from rick_b200 import _lib
    rng = np.random.default_rng(n)
    v = (rng.random(n) ** 4).astype(np.float32)
    if n > 8:
        v[3] = v[5] = v[7]                 # ties
        v[1] = -v[1]                       # a negative value
    qs = [40.0, 0.1, 75.0, 0.0, 100.0, 50.0, 33.3, 99.99]
    cv = torch.from_numpy(v).cuda()
    lines = torch.empty(8, dtype=torch.float64, device="cuda")
    q = (ctypes.c_double * 8)(*qs)
    _lib.check(_lib.lib().rick_percentile(lines.data_ptr(), cv.data_ptr(), n, q, 8,
                                          torch.cuda.current_stream().cuda_stream), "percentile")
    want = np.array([np.percentile(v.astype(np.float64), x) for x in qs])
    assert np.array_equal(lines.cpu().numpy(), want)


def test_near_tie_thresholds():
    """Adversarial: many filters within 1 ulp of the threshold -- decisions must still match NumPy exactly."""
    from rick_b200 import rick
    rng = np.random.default_rng(7)
    fg = synth.fisher_g(55, size=32)
    base = np.float32(3.1e-7)
    for key in list(fg)[:3]:
        if fg[key].ndim == 5:
            v = np.full(fg[key].shape, base, np.float32)
            v[0, ::3] = np.nextafter(base, np.float32(1))
            v[0, 1::3] = np.nextafter(base, np.float32(0))
            fg[key] = v
    cg = _to_cuda(fg)
    masks = rick.FilterMasks(rick.generator_layers(cg), "cuda")
    masks.update(cg, 40.0, 0.1)
    fr, ft, pr, lines = ro.decide_g(fg, 40.0, 0.1, n_convs=6)
    gfr, gft, gpr, gzero = masks.index_sets()
    for k in fr:
        assert np.array_equal(gfr[k], fr[k]) and np.array_equal(gft[k], ft[k]) and np.array_equal(gpr[k], pr[k]), k


def test_numpy1_comparison_rule_flag():
    """The reference pins NumPy 1.23.1, where ``float32_array > float64_scalar`` compares in float32 (value-based casting);
    NumPy >= 2 compares in float64.  ``FilterMasks(numpy1_compare=True)`` / RICK_DECIDE_COMPARE_F32 restate the former.
    Built so the rules disagree: an interpolated percentile lies strictly between two adjacent float32 FIM values, so in
    float64 one value is above the line and in float32 (line rounded onto one of the two) it is not."""
    from rick_b200 import rick
    fg = synth.fisher_g(55, size=32)
    cg = _to_cuda(fg)
    res = {}
    for flag in (False, True):
        masks = rick.FilterMasks(rick.generator_layers(cg), "cuda", numpy1_compare=flag)
        masks.update(cg, 40.0, 0.1)
        ro.NUMPY1_COMPARE = flag
        try:
            fr, ft, pr, lines = ro.decide_g(fg, 40.0, 0.1, n_convs=6)
        finally:
            ro.NUMPY1_COMPARE = False
        gfr, gft, gpr, _ = masks.index_sets()
        for k in fr:
            assert np.array_equal(gfr[k], fr[k]) and np.array_equal(gft[k], ft[k]) and np.array_equal(gpr[k], pr[k]), (flag, k)
        res[flag] = (gfr, gpr)
    # direct kernel check on a constructed disagreement
    from rick_b200 import _lib
    lo = np.float32(1.0)
    hi = np.nextafter(lo, np.float32(2))
    line = float(lo) + 0.75 * (float(hi) - float(lo))          # rounds to hi in float32
    fim = torch.tensor([lo, hi], dtype=torch.float32, device="cuda")
    lines = torch.tensor([line, -1.0], dtype=torch.float64, device="cuda")
    out = {}
    for flags in (0, 2):
        st = torch.zeros(2, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib().rick_decide(st.data_ptr(), None, fim.data_ptr(), 2, lines.data_ptr(), flags, 1,
                                          torch.cuda.current_stream().cuda_stream), "rick_decide")
        out[flags] = (st.cpu().numpy() & 1).tolist()
    assert out[0] == [0, 1]           # float64: hi > line
    assert out[2] == [0, 0]           # float32: hi > float32(line) == hi is False


def test_fisher_accumulate_matches_numpy_float32_order():
    from rick_b200 import rick
    g = torch.Generator().manual_seed(3)
    shapes = [(1, 8, 4, 3, 3), (16, 512), (16,), (1,), (3, 5, 7), (1000003,)]
    params = [(f"p{i}", torch.zeros(s, device="cuda")) for i, s in enumerate(shapes)]
    acc = rick.FisherAccumulator(params)
    total = {}
    for img in range(5):
        grads = [torch.randn(s, generator=g) * 1e-3 for s in shapes]
        acc.add([t.cuda() for t in grads])
        ro.fisher_accumulate(total, {f"p{i}": (t ** 2).numpy() for i, t in enumerate(grads)})
    acc.average(10)
    ro.fisher_average(total, 5, 2)
    for k, v in acc.as_dict().items():
        assert np.array_equal(v.cpu().numpy(), total[k]), k


def test_mask_apply_matches_reference_indexing():
    from rick_b200 import rick
    g = torch.Generator().manual_seed(4)
    fg = synth.fisher_g(56, size=32)
    cg = _to_cuda(fg)
    masks = rick.FilterMasks(rick.generator_layers(cg), "cuda")
    masks.update(cg, 40.0, 5.0)                # prune 5 % so the zero set is non-trivial
    fr, ft, pr, _ = ro.decide_g(fg, 40.0, 5.0, n_convs=6)
    params, np_params, np_grads = {}, {}, {}
    for k, v in fg.items():
        p = torch.nn.Parameter(torch.randn(v.shape, generator=g).cuda())
        p.grad = torch.randn(v.shape, generator=g).cuda()
        params[k] = p
        np_params[k], np_grads[k] = p.detach().cpu().numpy().copy(), p.grad.cpu().numpy().copy()
    masks.apply(params)
    ro.apply_masks_numpy(np_params, np_grads, fr, pr)
    for k in params:
        assert np.array_equal(params[k].detach().cpu().numpy(), np_params[k]), k
        assert np.array_equal(params[k].grad.cpu().numpy(), np_grads[k]), k


# ------------------------------------------------------------------------------------------ fused optimiser step
def _adam_case(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = {"convs.0.conv.weight": (1, 12, 8, 3, 3), "convs.0.conv.modulation.weight": (8, 16),
              "convs.0.conv.modulation.bias": (8,), "convs.0.activate.bias": (12,), "style.1.weight": (16, 16),
              "odd.weight": (5, 7)}
    return {n: torch.randn(s, generator=g) for n, s in shapes.items()}


def test_fused_masked_adam_matches_masks_plus_torch_adam_plus_ema():
    """rick_adam_mask_ema == rick_mask_apply semantics + torch.optim.Adam + accumulate(), over several steps."""
    from rick_b200.optim import FusedMaskedAdam

    class _Masks:
        def __init__(self, state, zero):
            self.state, self.zero = state, zero

    init = _adam_case()
    trainable = [n for n in init if n.startswith("convs") or n.startswith("odd")]
    gen = torch.Generator().manual_seed(1)
    state = {"convs.0.conv.weight": torch.tensor([1, 0, 4, 3, 0, 1, 4, 4, 0, 2, 1, 0], dtype=torch.uint8),
             "convs.0.conv.modulation.weight": torch.tensor([0, 1, 3, 4, 0, 0, 1, 4], dtype=torch.uint8)}
    state["convs.0.conv.modulation.bias"] = state["convs.0.conv.modulation.weight"].clone()
    zero = {k: ((v & 2) > 0).to(torch.uint8) for k, v in state.items()}
    zero["convs.0.conv.weight"][4] = 1                       # pruned in an earlier round, not in this one
    decay, lr, betas = 0.5 ** (32 / 10000), 0.002 * 4 / 5, (0.0, 0.99 ** (4 / 5))

    # --- reference composition on CPU (float32, torch.optim.Adam single-tensor path)
    ref = {n: torch.nn.Parameter(v.clone()) for n, v in init.items()}
    ref_ema = {n: v.clone() + 0.25 for n, v in init.items()}
    opt = torch.optim.Adam([ref[n] for n in trainable], lr=lr, betas=betas)
    # --- fused path
    dev = {n: torch.nn.Parameter(v.clone().cuda()) for n, v in init.items()}
    dev_ema = {n: torch.nn.Parameter((v.clone() + 0.25).cuda()) for n, v in init.items()}
    masks = _Masks({k: v.cuda() for k, v in state.items()}, {k: v.cuda() for k, v in zero.items()})
    fopt = FusedMaskedAdam(dev, [dev[n] for n in trainable], lr=lr, betas=betas, masks=masks, ema_named=dev_ema,
                           ema_decay=decay)
    for step in range(4):
        grads = {n: torch.randn(init[n].shape, generator=gen) * 10 ** (-step) for n in trainable}
        use_masks = step > 0
        for n in trainable:
            ref[n].grad = grads[n].clone()
            dev[n].grad = grads[n].clone().cuda()
        if use_masks:
            with torch.no_grad():
                for n, st in state.items():
                    p = ref[n]
                    rows = p.view(st.numel(), -1)
                    grows = p.grad.view(st.numel(), -1)
                    z = zero[n].bool()
                    grows[(st & 1).bool() | z] = 0
                    rows[z] = 0
        opt.step()
        with torch.no_grad():
            for n in init:
                ref_ema[n].mul_(decay).add_(ref[n].detach(), alpha=1 - decay)
        fopt.step(apply_masks=use_masks, ema=True)
    for n in init:
        torch.testing.assert_close(dev[n].detach().cpu(), ref[n].detach(), rtol=2e-6, atol=1e-7, msg=n)
        torch.testing.assert_close(dev_ema[n].detach().cpu(), ref_ema[n], rtol=2e-6, atol=1e-7, msg=n)
    # masks are exact: pruned filters are exactly zero, frozen filters did not move after masking started
    w = dev["convs.0.conv.weight"].detach().cpu().view(12, -1)
    assert (w[zero["convs.0.conv.weight"].bool()] == 0).all()
    # untouched (non-trainable) parameter: bit-identical
    assert torch.equal(dev["style.1.weight"].detach().cpu(), init["style.1.weight"])
    # optimiser state moves into torch.optim.Adam unchanged
    sd = fopt.state_dict()
    opt2 = torch.optim.Adam([torch.nn.Parameter(init[n].clone().cuda()) for n in trainable], lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == pytest.approx(lr) and float(opt2.state_dict()["state"][0]["step"]) == 4
    fopt2 = FusedMaskedAdam(dev, [dev[n] for n in trainable], lr=1.0, betas=(0.5, 0.5))
    fopt2.load_state_dict(opt.state_dict())
    assert fopt2.lr == pytest.approx(lr) and fopt2.steps.tolist() == [4.0] * len(trainable)
    torch.testing.assert_close(fopt2.exp_avg_sq[id(dev[trainable[0]])].cpu(),
                               opt.state_dict()["state"][0]["exp_avg_sq"], rtol=0, atol=0)


def test_fused_masked_adam_ema_only_and_channels_last():
    from rick_b200.optim import FusedMaskedAdam
    g = torch.Generator().manual_seed(3)
    w = torch.randn(16, 8, 3, 3, generator=g)
    p = torch.nn.Parameter(w.cuda().to(memory_format=torch.channels_last))
    e = torch.nn.Parameter((w * 0.5).cuda().to(memory_format=torch.channels_last))
    opt = FusedMaskedAdam({"convs.1.0.weight": p}, [p], lr=0.01, betas=(0.0, 0.99), ema_named={"convs.1.0.weight": e},
                          ema_decay=0.9)
    opt.ema_only()
    torch.testing.assert_close(e.detach().cpu(), w * 0.5 * 0.9 + w * 0.1, rtol=1e-6, atol=1e-7)
    assert float(opt.steps.sum()) == 0 and torch.equal(p.detach().cpu(), w)
    p.grad = torch.randn(16, 8, 3, 3, generator=g).cuda()          # contiguous gradient for a channels-last weight
    with pytest.raises(RuntimeError, match="memory layout"):
        opt.step()
