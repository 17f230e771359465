"""GPU parity of the pointwise companions (SRTM / LFGA / TEPD / the Sample.x square hook) against the CPU oracle.

The kernels compute in fp32 with separate roundings for every storage format, so the bar is bit-exactness:
  RGBA32F          : identical bits to the oracle (= the reference's F functions, tests/test_oracle.py)
  RGBA16F          : identical bits to round-to-nearest-even-half(oracle(half input as float))
  UNORM in / out   : identical integers to quantise(oracle(dequantise(input))) with the D3D conversions
"""
import numpy as np
import pytest
import torch

import fsr1_b200 as F
import oracle_lib as ol

pytestmark = pytest.mark.gpu
api = F.api
SIZES = [(96, 54), (33, 17), (5, 4), (300, 9), (257, 130)]


def _hdr(w, h, seed):
    a = F.structured(w, h, seed).copy()
    a[..., :3] = a[..., :3] ** 3 * 60.0
    a[::7, ::5, :3] = 0.0
    a[3::11, 2::3, :3] = 1.0
    return a


def _run(fn, src, out_like=None, **kw):
    d = torch.from_numpy(np.ascontiguousarray(src)).cuda()
    o = torch.zeros_like(d) if out_like is None else out_like
    fn(d, o, **kw)
    torch.cuda.synchronize()
    return o.cpu().numpy()


def _bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint16)


def _to_half_exact(a):
    """fp32 result -> what an RGBA16F store of it holds (round to nearest even; overflow -> inf like the hardware)."""
    with np.errstate(over="ignore"):
        return a.astype(np.float16)


@pytest.mark.parametrize("size", SIZES)
def test_srtm_and_inverse(size):
    w, h = size
    hdr = _hdr(w, h, 6)
    for inverse, src in ((False, hdr), (True, ol.srtm(hdr))):
        want = ol.srtm(src, inverse=inverse)
        got = _run(lambda a, b: api.srtm(a, b, inverse=inverse), src)
        assert api.last_kernel() == ("pointwise<srtm_inv>" if inverse else "pointwise<srtm>")
        assert np.array_equal(_bits(got), _bits(want))
        # fp16 storage: fp32 arithmetic on the half values, one rounding on store
        sh = F.to_half(src)
        wanth = _to_half_exact(ol.srtm(sh.astype(np.float32), inverse=inverse))
        goth = _run(lambda a, b: api.srtm(a, b, inverse=inverse), sh)
        assert np.array_equal(_bits(goth), _bits(wanth))
    # in place, on a row range only: rows outside [y0,y1) keep their content
    d = torch.from_numpy(hdr).cuda()
    y0, y1 = h // 3, max(h // 3 + 1, 2 * h // 3)
    api.srtm(d, d, y0=y0, y1=y1)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    assert np.array_equal(_bits(got[y0:y1]), _bits(ol.srtm(hdr)[y0:y1]))
    assert np.array_equal(_bits(got[:y0]), _bits(hdr[:y0])) and np.array_equal(_bits(got[y1:]), _bits(hdr[y1:]))


@pytest.mark.parametrize("size", SIZES)
def test_lfga(size):
    w, h = size
    img = F.structured(w, h, 5)
    img[0, 0, :3] = (0.0, 1.0, 0.5)
    grain = (F.uniform(16, 8, 77) - 0.5).astype(np.float32)
    dg = torch.from_numpy(grain).cuda()
    for amount in (0.0, 0.35, 1.0):
        want = ol.lfga(img, grain, amount)
        got = _run(lambda a, b: api.lfga(a, dg, b, amount), img)
        assert api.last_kernel() == "pointwise<lfga>"
        assert np.array_equal(_bits(got), _bits(want))
        assert np.array_equal(got[..., 3], img[..., 3])
    # half image + half grain tile
    ih, gh = F.to_half(img), F.to_half(grain)
    dgh = torch.from_numpy(gh).cuda()
    want = _to_half_exact(ol.lfga(ih.astype(np.float32), gh.astype(np.float32), 0.5))
    got = _run(lambda a, b: api.lfga(a, dgh, b, 0.5), ih)
    assert np.array_equal(_bits(got), _bits(want))


def _q(x, n):
    s = np.float32((1 << n) - 1)
    return (np.nan_to_num(np.clip(x, 0.0, 1.0), nan=0.0).astype(np.float32) * s + np.float32(0.5)).astype(np.uint32)


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("bits", [8, 10])
def test_tepd(size, bits):
    w, h = size
    img = F.structured(w, h, 5)
    noise = F.uniform(8, 8, 3)
    noise[0, 0, 3], noise[0, 1, 3] = -0.5, 1.5
    dn = torch.from_numpy(noise).cuda()
    for frame, dither, ddev in ((0, None, None), (77, None, None), (9, noise, dn)):
        want = ol.tepd(img, bits, frame=frame, dither=dither)
        got = _run(lambda a, b: api.tepd(a, b, bits, frame=frame, dither=ddev), img)
        assert api.last_kernel() == "pointwise<tepd%d>" % bits
        assert np.array_equal(_bits(got), _bits(want))
        # straight into the UNORM image the codes are meant for
        if bits == 8:
            out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        else:
            out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        api.tepd(torch.from_numpy(img).cuda(), out, bits, frame=frame, dither=ddev)
        torch.cuda.synchronize()
        a = out.cpu().numpy()
        if bits == 8:
            codes = a[..., :3].astype(np.uint32)
        else:
            u = a.view(np.uint32)
            codes = np.stack([u & 1023, (u >> 10) & 1023, (u >> 20) & 1023], axis=-1)
        assert np.array_equal(codes, _q(want[..., :3], bits))
    # half image: fp32 arithmetic on the half values; the result sits on a code value, rounded once to half
    ih = F.to_half(img)
    wanth = _to_half_exact(ol.tepd(ih.astype(np.float32), bits, frame=3))
    goth = _run(lambda a, b: api.tepd(a, b, bits, frame=3), ih)
    assert np.array_equal(_bits(goth), _bits(wanth))


def test_output_square_hook_on_the_last_pass():
    """FSR1_FLAG_OUTPUT_SQUARE = the sample's `if (Sample.x == 1) c *= c` (FSR_Pass.hlsl:78-79,93-94): applied to the
    output of the last pass only — RCAS when it runs, EASU with FSR1_FLAG_NO_RCAS."""
    iw, ih, ow, oh = 96, 54, 192, 108
    src = F.structured(iw, ih, 8)
    econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
    e = ol.easu(src, ow, oh)
    r = ol.rcas(e, ol.rcas_con(0.25))
    sq = lambda a: np.concatenate([(a[..., :3] * a[..., :3]).astype(np.float32), a[..., 3:]], axis=-1)
    d = torch.from_numpy(src).cuda()
    tmp, out = torch.zeros((oh, ow, 4), device="cuda"), torch.zeros((oh, ow, 4), device="cuda")
    api.upscale(d, tmp, out, econ, rcon, flags=api.FLAG_EXACT | api.FLAG_OUTPUT_SQUARE)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(sq(r)))
    assert np.array_equal(_bits(tmp.cpu().numpy()), _bits(e))                     # the intermediate stays un-squared
    api.upscale(d, tmp, out, econ, rcon, flags=api.FLAG_EXACT | api.FLAG_OUTPUT_SQUARE | api.FLAG_NO_RCAS)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(sq(e)))
    # fp16 production kernels: squared result within tolerance of the squared fp32 oracle, rows outside the slab untouched
    sh = F.to_half(src)
    eh = ol.easu(sh.astype(np.float32), ow, oh)
    rh = ol.rcas(eh, ol.rcas_con(0.25))
    dh = torch.from_numpy(sh).cuda()
    tmph = torch.zeros((oh, ow, 4), dtype=torch.float16, device="cuda")
    outh = torch.full((oh, ow, 4), 7.0, dtype=torch.float16, device="cuda")
    api.upscale(dh, tmph, outh, econ, rcon, y0=20, y1=60, flags=api.FLAG_OUTPUT_SQUARE)
    torch.cuda.synchronize()
    got = outh.cpu().numpy().astype(np.float32)
    assert np.abs(got[20:60, :, :3] - sq(rh)[20:60, :, :3]).max() <= 2e-2       # d(c^2) = 2c dc, c <= 1, dc <= 1e-2
    assert (got[:20] == 7.0).all() and (got[60:] == 7.0).all()


def test_sample_frame_chain_srtm_easu_rcas_inverse_lfga_tepd():
    """The order the passes compose in an application (ffx_fsr1.h:1030-1040 usage notes + the sample's tonemap->FSR
    chain): HDR -> SRTM -> EASU -> RCAS -> SRTM inverse -> (tone map, here identity on [0,1]) -> LFGA -> TEPD 10 bit.
    fp32 EXACT kernels: every stage bit-identical to the oracle chain."""
    iw, ih, ow, oh = 64, 36, 128, 72
    hdr = _hdr(iw, ih, 11)
    grain = (F.uniform(16, 16, 4) - 0.5).astype(np.float32)
    econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
    want = ol.srtm(hdr)
    want = ol.easu(want, ow, oh)
    want = ol.rcas(want, ol.rcas_con(0.25), True)
    want = ol.srtm(want, inverse=True)
    want = np.concatenate([np.clip(want[..., :3] * np.float32(1.0 / 64.0), 0, 1).astype(np.float32), want[..., 3:]], axis=-1)
    lf = ol.lfga(want, grain, 0.25)
    fin = ol.tepd(lf, 10, frame=5)
    d = torch.from_numpy(hdr).cuda()
    api.srtm(d, d)
    tmp, out = torch.zeros((oh, ow, 4), device="cuda"), torch.zeros((oh, ow, 4), device="cuda")
    api.upscale(d, tmp, out, econ, rcon, flags=api.FLAG_EXACT | api.FLAG_RCAS_CLAMP)
    api.srtm(out, out, inverse=True)
    out[..., :3] = torch.clamp(out[..., :3] * (1.0 / 64.0), 0, 1)
    api.lfga(out, torch.from_numpy(grain).cuda(), out, 0.25)
    code = torch.zeros((oh, ow), dtype=torch.int32, device="cuda")
    api.tepd(out, code, 10, frame=5)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(lf))
    u = code.cpu().numpy().view(np.uint32)
    codes = np.stack([u & 1023, (u >> 10) & 1023, (u >> 20) & 1023], axis=-1)
    assert np.array_equal(codes, _q(fin[..., :3], 10))


def test_fsr_filter_hdr_flag_squares_the_last_pass():
    """FSR_Filter::Upscale(..., hdr) (FSR_Filter.cpp:107,125): Sample.x = 1 on the pass that writes the output."""
    iw, ih, ow, oh = 96, 54, 192, 108
    src = torch.from_numpy(F.to_half(F.structured(iw, ih, 3))).cuda()
    flt = F.FSR_Filter()
    flt.OnCreate()
    flt.OnCreateWindowSizeDependentResources(iw, ih, ow, oh)
    for use_rcas in (True, False):
        st = F.State(renderWidth=iw, renderHeight=ih, rcasAttenuation=0.25, bUseRcas=use_rcas)
        plain = torch.zeros((oh, ow, 4), dtype=torch.float16, device="cuda")
        hdr = torch.zeros_like(plain)
        flt.Upscale(src, plain, ow, oh, st)
        flt.Upscale(src, hdr, ow, oh, st, hdr=True)
        torch.cuda.synchronize()
        want = plain.float()
        want[..., :3] = want[..., :3] * want[..., :3]
        assert torch.equal(hdr, want.half())
    flt.OnDestroy()


def _golden_runners_gpu():
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def srtm(img, inverse):
        return _run(lambda a, b: api.srtm(a, b, inverse=inverse), img)

    def lfga(img, grain, amount):
        g = up(grain)
        return _run(lambda a, b: api.lfga(a, g, b, amount), img)

    def tepd(img, bits, frame, dither):
        d = up(dither) if dither is not None else None
        return _run(lambda a, b: api.tepd(a, b, bits, frame=frame, dither=d), img)

    return srtm, lfga, tepd


def check_against_pointwise_golden(srtm, lfga, tepd):
    """Shared by the GPU test below and by tests/test_oracle.py-style CPU use: every committed array, bit for bit."""
    import os
    P = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fsr1_pointwise_golden.npz"))
    sdr, hdr, grain, noise = P["sdr"], P["hdr"], P["grain"], P["noise"]
    assert np.array_equal(_bits(srtm(hdr, False)), _bits(P["srtm"]))
    assert np.array_equal(_bits(srtm(P["srtm"], True)), _bits(P["srtm_inv"]))
    for amount in (0.0, 0.35, 1.0):
        assert np.array_equal(_bits(lfga(sdr, grain, amount)), _bits(P["lfga_%g" % amount])), amount
    for b in (8, 10):
        assert np.array_equal(_bits(tepd(sdr, b, 5, None)), _bits(P["tepd%d_f5" % b])), b
        assert np.array_equal(_bits(tepd(sdr, b, 0, noise)), _bits(P["tepd%d_noise" % b])), b


def test_pointwise_against_committed_golden_vectors():
    """The reference's own outputs (tests/golden/make_pointwise_golden.py), 300 px wide with 12x5 / 9x7 aux tiles: more
    than one pixel per thread in a row and a tile width that does not divide the thread stride."""
    check_against_pointwise_golden(*_golden_runners_gpu())
