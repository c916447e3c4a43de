"""CPU-only tests of the host side: the C ABI library loads and exports every symbol the
header declares, the nn.Module surface matches the reference's, host-side RNG decisions follow
the reference's call order, errors are loud, and the multi-rank plumbing works on gloo."""
import inspect
import os
import random
import re

import numpy as np
import pytest
import torch

import wmattack
from oracle import attack_oracle as O
from tests.golden_util import GOLD, T, text
from wmattack import _lib, modules, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from wmattack import build
        build.build()
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "wm_attack.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(?:int64_t|int|const char\s*\*)\s*(wm_\w+)\s*\(", src))


def test_library_exports_every_header_symbol(lib):
    names = header_functions()
    assert {"wm_version", "wm_last_error", "wm_diffjpeg_fwd", "wm_diffjpeg_bwd"} <= names
    for n in names:
        assert hasattr(lib, n), f"libwmattack.so does not export {n}"
    assert set(_lib.SIGNATURES) | set(_lib.HELPERS) == names - {"wm_version", "wm_last_error"}
    assert lib.wm_version() == 4
    assert isinstance(lib.wm_last_error(), bytes)


def header_prototypes():
    """name -> list of C parameter declarations, parsed from include/wm_attack.h."""
    src = open(os.path.join(ROOT, "include", "wm_attack.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int64_t|int|const char\s*\*)\s*(wm_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = [q.strip() for q in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params == ["void"] else params
    return out


def test_ctypes_signatures_match_the_header_parameter_by_parameter():
    """The ctypes table is written by hand: every entry point must have as many arguments as its prototype, and each
    argument the right CLASS (pointer / 64-bit integer / 32-bit integer / float) - a swapped or missing argument would
    otherwise only show up as garbage on the GPU."""
    import ctypes as C
    protos = header_prototypes()
    table = dict(_lib.SIGNATURES)
    table.update({k: v[1] for k, v in _lib.HELPERS.items()})

    def c_class(decl):
        if "*" in decl:
            return "ptr"
        t = decl.replace("const", "").split()
        if t[0] in ("int64_t", "uint64_t"):
            return "i64"
        if t[0] in ("int", "unsigned", "uint32_t", "int32_t"):
            return "i32"
        if t[0] == "float":
            return "f32"
        raise AssertionError(f"unparsed parameter {decl!r}")

    def ct_class(t):
        if t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_int))) or getattr(t, "_type_", None) == "P":
            return "ptr"
        if t in (C.c_int64, C.c_uint64, C.c_size_t):
            return "i64"
        if t in (C.c_int, C.c_uint):
            return "i32"
        if t is C.c_float:
            return "f32"
        raise AssertionError(f"unclassified ctypes type {t}")

    for name, argtypes in table.items():
        assert name in protos, name
        want = [c_class(d) for d in protos[name]]
        got = [ct_class(t) for t in argtypes]
        assert got == want, f"{name}: ctypes {got} vs header {want}"


def test_invalid_arguments_fail_loudly_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so it is testable on a CPU box
    with pytest.raises(_lib.WMAttackError, match="null"):
        _lib.call("wm_diffjpeg_fwd", None, 0, 0, 0, 0, None, 1, 32, 32, 1.0, None, 0, None, None)
    with pytest.raises(_lib.WMAttackError, match="multiples of 16"):
        _lib.call("wm_diffjpeg_fwd", 256, 0, 8, 8, 8, 256, 1, 24, 24, 1.0, None, 0, None, None)
    with pytest.raises(_lib.WMAttackError, match="aligned"):
        _lib.call("wm_diffjpeg_fwd", 260, 0, 8, 8, 8, 256, 1, 32, 32, 1.0, None, 0, None, None)
    with pytest.raises(_lib.WMAttackError, match="kernel size"):
        _lib.call("wm_median_fwd", 256, 0, 0, 256, None, 0, 1, 8, 8, 4, None, None)
    with pytest.raises(_lib.WMAttackError, match="odd"):
        taps = (_lib.f32 * 4)(0.25, 0.25, 0.25, 0.25)
        _lib.call("wm_gaussblur", 256, 64, 8, 256, 1, 8, 8, taps, 4, 0, 0, None, None)
    with pytest.raises(_lib.WMAttackError, match="mode"):
        _lib.call("wm_interp_fwd", 256, 64, 8, 8, 8, 0, 0, 8, 8, 256, 1, 4, 4, 7, 0, None, None)
    with pytest.raises(_lib.WMAttackError, match="outside"):       # crop window beyond the source plane
        _lib.call("wm_interp_fwd", 256, 64, 8, 8, 8, 4, 0, 8, 8, 256, 1, 4, 4, 0, 0, None, None)
    # the store epilogue is an explicit argument: a misaligned x, or a code path that cannot apply it, is an error
    ep = _lib.StoreEpilogue(260, 1, 1)
    import ctypes
    with pytest.raises(_lib.WMAttackError, match="epilogue"):
        _lib.call("wm_diffjpeg_fwd", 256, 0, 3072, 1024, 32, 256, 1, 32, 32, 1.0, None, 0, ctypes.byref(ep), None)
    ep = _lib.StoreEpilogue(256, 1, 1)
    with pytest.raises(_lib.WMAttackError, match="does not apply"):
        taps = (_lib.f32 * 9)(*([1 / 9] * 9))                      # k = 9: generic blur path
        _lib.call("wm_gaussblur", 256, 64, 8, 256, 1, 8, 8, taps, 9, 0, 0, ctypes.byref(ep), None)
    assert lib.wm_interp_is_tiled(512, 512, 256, 256, 192) == 1 and lib.wm_interp_is_tiled(512, 512, 64, 64, 3) == 0


def test_no_cpu_fallback():
    x = torch.rand(1, 3, 32, 32)
    for layer in (wmattack.DiffJPEG(True, 32, 32, 50), wmattack.Jpeg(50), wmattack.JpegCompression("cpu"),
                  wmattack.GaussianBlur(), wmattack.MiddleBlur(3), wmattack.Gaussian(), wmattack.SaltPepper(0.1),
                  wmattack.Resize(), wmattack.Quantization()):
        with pytest.raises(RuntimeError, match="CUDA"):
            layer(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        wmattack.Crop()(x)
    # the product never imports the oracle
    for mod in (modules, wmattack.functional, _lib):
        assert "oracle" not in open(inspect.getsourcefile(mod)).read().replace("the oracle", "")


def test_module_surface_matches_reference():
    sig = lambda f: str(inspect.signature(f))
    assert sig(wmattack.DiffJPEG.__init__) == "(self, differentiable=True, height=512, width=512, quality=75, rounding=<function round_only_at_0>)".replace(
        "<function round_only_at_0>", repr(modules.round_only_at_0))
    assert sig(wmattack.DiffJPEG.forward) == "(self, image, quality=None)"
    assert sig(wmattack.Jpeg.__init__) == "(self, Q, subsample=0)"
    assert sig(wmattack.JpegCompression.__init__) == "(self, device=None, yuv_keep_weights=(25, 9, 9))"
    assert sig(wmattack.GaussianBlur.__init__) == "(self, kernel_size=3, channels=3)"
    assert sig(wmattack.GaussianBlur.forward) == "(self, tensor, cover_image=None)"
    assert sig(wmattack.Gaussian.forward) == "(self, tensor, cover_image=None, mean=0, stddev=0.05, noise=None)"
    assert sig(wmattack.Resize.__init__) == "(self, resize_ratio_range=(0.5, 1.5), interpolation_method='bicubic')"
    assert sig(wmattack.Resize.forward) == "(self, noised_image, resize_ratio=None)"
    assert sig(wmattack.Crop.forward) == "(self, image, apex=None, min_rate=0.5, max_rate=1.0)"
    assert sig(wmattack.Crop.cropped_out) == "(self, image, apex=None, min_rate=None, max_rate=1.0)"
    assert sig(wmattack.Combined.forward) == "(self, image_and_cover, id=None)"
    assert sig(wmattack.MaskDropout.forward).startswith("(self, noised_image, cover_image")
    names = {
        "Identity": wmattack.Identity().name, "Jpeg50": wmattack.Jpeg(50).name, "JpegSS70": wmattack.JpegSS(70).name,
        "JpegMask30": wmattack.JpegMask(30).name, "DiffJPEG50": wmattack.DiffJPEG(True, 32, 32, 50).name,
        "DiffJPEG75": wmattack.DiffJPEG(90).name, "Resize": wmattack.Resize().name, "G_Blur": wmattack.GaussianBlur().name,
        "Gaussian": wmattack.Gaussian().name, "Dropout": wmattack.MaskDropout().name,
        "NotChosenYet": wmattack.Combined().name, "MiddleBlur5": wmattack.MiddleBlur(5).name,
    }
    for want, got in names.items():
        assert want == got
    assert text("jpeg/name_q50") == wmattack.Jpeg(50).name and text("diffjpeg/name_q30") == wmattack.DiffJPEG(True, 48, 32, 30).name
    # drop-in overlay import paths of the reference (models/IRN_model.py:6,19,31-38)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200", "dropin"))
    try:
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k.startswith("noise_layers")]:
            del sys.modules[m]
        import noise_layers
        from noise_layers import (Identity, Crop, Cropout, Dropout, GN, MiddleBlur, GF, SaltPepper, Jpeg, JpegSS,  # noqa: F401
                                  JpegMask, JpegTest, Combined, get_random_int, get_random_float)
        from noise_layers.dropout import Dropout as D2
        from noise_layers.gaussian import Gaussian  # noqa: F401
        from noise_layers.gaussian_blur import GaussianBlur  # noqa: F401
        from noise_layers.middle_filter import MiddleBlur as M2  # noqa: F401
        from noise_layers.resize import Resize  # noqa: F401
        from noise_layers.jpeg_compression import JpegCompression  # noqa: F401
        from noise_layers.salt_pepper_noise import SaltPepper as S2  # noqa: F401
        from utils.JPEG import DiffJPEG, round_only_at_0, diff_round, quality_to_factor  # noqa: F401
        assert Dropout is wmattack.ElementDropout and D2 is wmattack.MaskDropout
        assert noise_layers.Combined is wmattack.Combined and DiffJPEG is wmattack.DiffJPEG
    finally:
        sys.path.pop(0)


def test_host_constants_match_oracle():
    assert wmattack.quality_to_factor(50) == 1.0 and wmattack.quality_to_factor(10) == 5.0
    assert wmattack.quality_to_factor(75) == O.quality_to_factor(75)
    for q in (10, 30, 50, 75, 90, 95):
        m = wmattack.Jpeg(q)
        ly, lc = O.jpeg8_tables(O.jpeg8_scale(q))
        tab = np.array([[m._params.table[c][i] for i in range(64)] for c in range(3)]).reshape(3, 8, 8)
        assert np.array_equal(tab[0], ly.numpy()) and np.array_equal(tab[1], lc.numpy()) and np.array_equal(tab[2], lc.numpy())
        assert tab.min() >= 1
    jc = wmattack.JpegCompression(None)
    for c, k in enumerate((25, 9, 9)):
        tab = np.array([jc._params.table[c][i] for i in range(64)]).reshape(8, 8)
        assert np.array_equal(tab, O.zigzag_keep_mask(k)) and tab.sum() == k
    for k in (3, 5, 7):
        assert np.allclose(wmattack.GaussianBlur(k)._taps, O.gaussian_taps(k).numpy(), atol=1e-15)
    assert np.allclose(wmattack.GaussianBlur(3)._taps, [0.3192, 0.3616, 0.3192], atol=1e-4)
    from wmattack.modules import _rounding_mode, diff_round, round_only_at_0
    assert _rounding_mode(round_only_at_0) == 0 and _rounding_mode(diff_round) == 1 and _rounding_mode(torch.round) == 2
    x = torch.linspace(-2, 2, 101, dtype=torch.float64)
    assert torch.equal(round_only_at_0(x), O.apply_rounding(x, O.ROUND_ONLY_AT_0))
    assert torch.equal(diff_round(x), O.apply_rounding(x, O.ROUND_CUBIC))


def test_host_rng_decisions_follow_the_reference():
    # Crop: rand() x2 -> coupling -> randint per free axis (noise_layers/crop.py:13-40)
    c = wmattack.Crop()
    np.random.seed(18)
    c._ratios(0.5, 1.0, 0.2)
    box = c.get_random_rectangle_inside((2, 3, 32, 32), c.height_ratio, c.width_ratio)
    assert tuple(box) == tuple(int(v) for v in GOLD["crop/seed18/x32/apex"])
    assert tuple(box) == O.crop_box_from_rng((2, 3, 32, 32), np.random.RandomState(18))
    # Resize ratio: np.random.rand() (noise_layers/resize.py:13)
    np.random.seed(17)
    assert modules.random_float(0.5, 1.5) == float(GOLD["resize/random/ratio"])
    # Combined: python random.randint (noise_layers/__init__.py:8-9); Identity members run on CPU
    class Named(wmattack.Identity):
        def __init__(self, n):
            super().__init__()
            self.name = n
    random.seed(21)
    comb = wmattack.Combined([Named("Identity"), Named("Jpeg50"), Named("JpegSS70"), Named("JpegMask30"), Named("Resize")])
    got = []
    for _ in range(12):
        comb(torch.zeros(1))
        got.append(comb.name)
    assert ",".join(got) == text("combined/seed21/names")
    assert comb(torch.zeros(1), id=2) is not None and comb.name == "JpegSS70"
    assert comb(torch.zeros(1), id=99) is not None          # out-of-range id -> random pick, as upstream
    x = torch.zeros(2)
    assert wmattack.Identity()(x) is x and int(GOLD["identity/is_same_object"]) == 1


def test_frame_sharding():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    x = torch.arange(10).view(10, 1)
    assert sharding.shard_batch(x, 1, 3).flatten().tolist() == [4, 5, 6]
    with pytest.raises(ValueError):
        sharding.frame_shard(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        from wmattack import functional as WF
        from wmattack.sharding import frame_shard
        torch.manual_seed(10)        # every rank seeds alike, as train.py:317-329 does
        span = frame_shard(13, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, span)
        seed, off = WF.next_philox_stream(1000)
        offs = [None] * world
        dist.all_gather_object(offs, (seed, off))
        mx = bench.allreduce_max(10.0 + rank, world, torch.device("cpu"))
        bench.barrier(world)
        q.put((rank, spans, offs, mx))
    finally:
        dist.destroy_process_group()


def test_two_rank_plumbing_on_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, spans, offs, mx in res:
        assert spans == [(0, 7), (7, 13)]
        assert offs[0][0] == offs[1][0] and offs[0][1] != offs[1][1]      # same seed, disjoint Philox sub-streams
        assert mx == 11.0                                                  # max over ranks


def test_resize_band_bound():
    """The fused Resize kernel keeps, per lane, a register window of 8/10/12/14 taps of the banded
    operator A = U D (and of A^T).  Mirror of csrc/resize_fused.cu::rb_band: the window must hold the
    band made regular over the outputs of a warp (H pass: two outputs per lane share a 10- or 14-wide window) and
    the band shared by a quad of output rows (V pass: window + 2)."""
    import math
    import numpy as np
    from oracle import attack_oracle as O

    def window(n, nm):
        fl = math.floor(3.0 * np.float32(n) / np.float32(nm) + 1e-3)
        need = 8 if fl <= 2 else fl + 7
        return 8 if need <= 8 else 10 if need <= 10 else 12 if need <= 12 else 14

    for n in (20, 64, 100, 512):
        for ratio in (0.45, 0.5, 0.53, 0.6, 0.667, 0.75, 0.8, 0.99, 1.0, 1.01, 1.25, 1.5, 2.0, 2.2):
            nm = int(ratio * n)
            if nm < 1 or not (0.45 <= nm / n <= 2.2):
                continue
            for mode in ("bicubic", "bilinear"):
                A = (O.interp_matrix(nm, n, mode) @ O.interp_matrix(n, nm, mode)).numpy()
                for M in (A, A.T):
                    nz = [np.nonzero(r)[0] for r in M]      # a source index bilinear never samples has an empty row of A^T
                    keep = [k for k, i in enumerate(nz) if len(i)]
                    if len(keep) < len(nz):                 # empty rows: zero weights, any in-range start works
                        for k in range(len(nz)):
                            if not len(nz[k]):
                                nb = min(keep, key=lambda j: abs(j - k))
                                nz[k] = np.array([min(max(nz[nb][0] + (k - nb), 0), n - 1)])
                    lo = np.array([i[0] for i in nz]); hi = np.array([i[-1] for i in nz])
                    need = 0
                    for g0 in range(0, n, 32):            # H pass: regular starts base + lane
                        l, h = lo[g0:g0 + 32], hi[g0:g0 + 32]
                        base = (l - np.arange(len(l))).min()
                        need = max(need, int((h - (base + np.arange(len(l))) + 1).max()))
                    need2 = 0
                    for g0 in range(0, n, 64):            # H pass as built: two outputs per lane, even regular starts
                        l, h = lo[g0:g0 + 64], hi[g0:g0 + 64]
                        if len(l) % 2:
                            l, h = np.append(l, l[-1]), np.append(h, h[-1])
                        lane = np.arange(len(l) // 2)
                        base = int((l[0::2] - 2 * lane).min()); base -= base % 2
                        need2 = max(need2, int((np.maximum(h[0::2], h[1::2]) - (base + 2 * lane) + 1).max()))
                    quad = max(int(hi[o:o + 4].max() - lo[o] + 1) for o in range(0, n, 4))   # V pass: 4 rows share a window
                    bt = window(n, nm)
                    assert need <= bt and need2 <= (10 if bt <= 10 else 14) and quad <= bt + 2, (n, nm, mode, need, need2, quad, bt)


# ---------------------------------------------------------------- generated median networks (csrc/*.cuh)
def _run_network(path, func, values):
    """Execute the min/max statements of a generated network header on a python list."""
    import re
    src = open(path).read()
    body = src[src.index(func + "("):]
    body = body[:body.index("\n}")]
    v = list(values)
    for ln in body.split("\n"):
        ln = ln.strip()
        m = re.match(r"(?:WM_CE|ce)\(v\[(\d+)\], v\[(\d+)\]\);", ln)
        if m:
            a, b = int(m[1]), int(m[2])
            v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])
            continue
        m = re.match(r"v\[(\d+)\] = (fminf|fmaxf)\(v\[(\d+)\], v\[(\d+)\]\);", ln)
        if m:
            f = min if m[2] == "fminf" else max
            v[int(m[1])] = f(v[int(m[3])], v[int(m[4])])
            continue
        m = re.match(r"return v\[(\d+)\];", ln)
        if m:
            return v[int(m[1])], v
    return None, v


def test_generated_median_networks_select_the_median():
    """The 5x5 kernel relies on generated selection networks; replay their statements on random
    inputs WITH TIES (the generator itself proves them with the 0-1 principle)."""
    import numpy as np
    csrc = os.path.join(ROOT, "video-watermarking-forgery-detection_b200", "csrc")
    rng = np.random.RandomState(0)
    for _ in range(300):
        rows = np.sort(rng.randint(0, 12, (6, 5)), axis=1)          # six sorted window rows, many ties
        _, v = _run_network(os.path.join(csrc, "median_pair_net.cuh"), "mid6_of_4_sorted_rows", rows[1:5].ravel())
        assert v[7:13] == list(np.sort(rows[1:5].ravel())[7:13])
        # the two-level form the kernel uses: merge row pairs, then the middle six of two merged pairs
        pairs = []
        for a, b in ((1, 2), (3, 4)):
            _, m = _run_network(os.path.join(csrc, "median_pair_net.cuh"), "merge10_sorted_5_5", list(rows[a]) + list(rows[b]))
            assert m == sorted(list(rows[a]) + list(rows[b]))
            pairs += m
        _, v2 = _run_network(os.path.join(csrc, "median_pair_net.cuh"), "mid6_of_2_sorted_10", pairs)
        assert v2[7:13] == v[7:13]
        for own, window in ((rows[0], rows[:5]), (rows[5], rows[1:])):
            med, _ = _run_network(os.path.join(csrc, "median_pair_net.cuh"), "median11_sorted_6_5", v[7:13] + list(own))
            assert med == np.sort(window.ravel())[12]
