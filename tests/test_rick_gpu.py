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
