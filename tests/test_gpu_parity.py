"""GPU parity tests: the CUDA path, called through the C ABI (via the host mirror), against
  (1) the golden fixtures produced by the unmodified reference (tests/golden/*.npz), and
  (2) the numpy oracle on seeded inputs.
Bar: bit-exact for indices, gathers with the triangular encoding, quantiser codes; rtol 1e-5 for the fp32 MLP path;
+-1 LSB on 8-bit output for >= 99.9 % of texels and PSNR within 0.05 dB for the f16/bf16 tensor-core path."""
import hashlib
import os

import numpy as np
import pytest
import torch

import inputs as I
from helpers import T, configure, dev, load, lsb_stats, make_decoder, psnr256
from oracle import nic_oracle as O

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def nic():
    import neural_image_compression_v2_b200 as n
    return n


# ------------------------------------------------------------------------------------------------ quantisers, PE
@pytest.mark.parametrize("bits", [8, 4, 2])
def test_quantisers_bit_exact(bits):
    m = nic().models
    z = load("quant.npz")
    x = z[f"x{bits}"]
    q_min, q_max = O.q_range(bits)
    ok = (x >= np.float32(q_min)) & (x <= np.float32(q_max))
    xt = T(x)
    assert np.array_equal(m.quantize4fp(xt, bits).cpu().numpy(), z[f"q{bits}"])
    assert np.array_equal(m.save4fp(xt, bits, torch.uint8).cpu().numpy()[ok], z[f"code{bits}"][ok])
    codes = T(z[f"code{bits}"][ok])
    assert np.array_equal(m.load4fp(codes, bits, torch.float32).cpu().numpy(), z[f"load{bits}"][ok])
    assert np.array_equal(m.load4fp(codes, bits, torch.uint8).cpu().numpy(), z[f"load{bits}"][ok])   # dtype bug not reproduced
    assert np.array_equal(m.quantize_clamp(T(x * np.float32(1.5)), bits).cpu().numpy(), z[f"clamp{bits}"])
    # idempotence and code round trip (size-independent properties)
    q = m.quantize4fp(xt, bits)
    assert torch.equal(m.quantize4fp(q, bits), q)
    c = m.save4fp(T(x[ok]), bits)
    assert torch.equal(m.save4fp(m.load4fp(c, bits), bits), c)


def test_output_quantiser_and_psnr():
    n = nic()
    z = load("quant.npz")
    ok = z["y"] <= 1.0                               # the fixture also holds (255.5/255), outside the image range
    y8 = n.models.output_to_u8(T(z["y"]), 8)
    assert np.array_equal(y8.cpu().numpy().astype(np.float32)[ok], z["y_to8"][ok])
    assert np.array_equal(n.models.quantize_to_bit(T(z["y"]), 8).cpu().numpy()[ok], z["y_to8"][ok])
    a, b = T(z["psnr_a"], torch.uint8), T(z["psnr_b"], torch.uint8)
    assert abs(n.utils.calculate_psnr(a, b) - float(z["psnr"])) < 1e-4


def test_positional_encodings():
    u = nic().utils
    z = load("pe.npz")
    for name in ("c14", "dy2", "dy3"):
        c = T(z[name])
        assert np.array_equal(u.triangular_positional_encoding(c, 6, dev(), torch.float32).cpu().numpy(), z[f"tri_{name}"])
        s = u.positional_encoding(tuple(c), 6, dev(), torch.float32).cpu().numpy()
        np.testing.assert_allclose(s, z[f"sin_{name}"], rtol=0, atol=2e-6)
    assert np.array_equal(u.triangular_positional_encoding(T(z["dy2"]), 4, dev(), torch.float32).cpu().numpy(), z["tri4_dy2"])


# ------------------------------------------------------------------------------------------------ gather
def test_gather_2d_golden_bit_exact():
    ic = nic().image_compression
    z = load("gather_decode_2d.npz")
    size = int(z["image_size"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=6)
    fp = [T(z[f"grid{i}"]) for i in range(4)]
    for mip in range(7):
        s, x, y = [int(v) for v in z[f"block_mip{mip}"]]
        X = ic.finally_decode_input_2d(fp, s, mip, x, y).cpu().numpy()
        assert np.array_equal(X, z[f"X_block_mip{mip}"]), mip
        full = ic.finally_decode_input_2d(fp, size >> mip, mip).cpu().numpy()
        assert sha(full) == str(z[f"X_full_sha_mip{mip}"]), mip
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=6, TF_USE_TRI_PE=False)
    Xs = ic.finally_decode_input_2d(fp, 16, 0, 40, 24).cpu().numpy()
    assert np.array_equal(Xs[:, :60], z["X_block_sin_mip0"][:, :60])
    np.testing.assert_allclose(Xs, z["X_block_sin_mip0"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(ic.finally_decode_input_2d(fp, 8, 3, 0, 0).cpu().numpy(), z["X_block_sin_mip3"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("method", [3, 4])
def test_gather_3d_golden(method):
    ic = nic().image_compression
    z = load(f"gather_decode_3d_m{method}.npz")
    size, cin = int(z["image_size"]), int(z["cin"])
    configure(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=3)
    fp = [T(a) for a in I.make_grids(size, 3, seed=31)]
    fin = ic.finally_decode_input_3d if method == 3 else ic.finally_decode_input_3d_v2
    cre = ic.create_decoder_input_3d if method == 3 else ic.create_decoder_input_3d_v2
    for mip in range(6):
        s, x, y, zz = [int(v) for v in z[f"block_mip{mip}"]]
        X = fin(fp, s, mip, x, y, zz).cpu().numpy()
        ref = z[f"X_block_mip{mip}"]
        if method == 3:
            assert np.array_equal(X, ref), mip
        else:
            assert np.array_equal(X[:, :60], ref[:, :60]) and np.array_equal(X[:, -1], ref[:, -1])
            np.testing.assert_allclose(X, ref, rtol=0, atol=2e-6)
    coord = T(z["train_coord"])
    X0 = cre(fp, coord, 2, 0, 0).cpu().numpy()
    X1 = cre(fp, coord // 2, 2, 0, 1).cpu().numpy()
    if method == 3:
        assert np.array_equal(X0, z["X_train_mip0"]) and np.array_equal(X1, z["X_train_mip1"])
    else:
        np.testing.assert_allclose(X0, z["X_train_mip0"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(X1, z["X_train_mip1"], rtol=0, atol=2e-6)


def test_gather_vs_oracle_larger_and_ragged():
    """Seeded inputs the fixtures do not cover: 256^2 grids, non-square blocks, several crops, edge origins."""
    n = nic()
    L = n._lib
    size = 256
    grids = I.make_grids(size, 2, seed=77)
    fp = [T(a) for a in grids]
    table = O.create_pyramid_mip_levels(size, size // 4)
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    for mip, s, origin in ((0, 256, (0, 0)), (0, 37, (219, 1)), (1, 128, (0, 0)), (2, 64, (0, 0)), (3, 32, (0, 0)), (5, 8, (0, 0))):
        X = n.image_compression.finally_decode_input_2d(fp, s, mip, *origin).cpu().numpy()
        assert np.array_equal(X, O.finally_decode_input(grids, s, mip, table, 1, origin)), (mip, s)
    # ragged block (5 x 7) through the raw ABI geometry
    import ctypes as C
    fl = 0
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], (5, 7), 1, -2, 0, 6, L.PE_TRIANGULAR, origin0=(240, 3))
    x = torch.empty((35, 73), dtype=torch.float32, device=dev())
    h = L.handle(dev())
    L.check(h, L.load_library().nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, L.ptr(x), L.DT_F32, L.stream_ptr(dev())))
    ref = O.decoder_input_one(grids[0], grids[1], (240, 3), 7, 0.25, 0, 1)     # 7x7 block, take the first 5 x-rows
    assert np.array_equal(x.cpu().numpy(), ref.reshape(7, 7, 73)[:5].reshape(35, 73))
    # empty input
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], (0, 7), 1, -2, 0, 6, L.PE_TRIANGULAR)
    L.check(h, L.load_library().nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, None, L.DT_F32, L.stream_ptr(dev())))
    # out-of-bounds host origin is an error, like the reference's IndexError
    with pytest.raises(n.NicError) as e:
        n.image_compression.finally_decode_input_2d(fp, 256, 0, 1, 0)
    assert e.value.status == -4
    # 16-bit outputs are the fp32 value rounded once
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], 64, 1, -2, 0, 6, L.PE_TRIANGULAR)
    x32 = torch.empty((4096, 73), dtype=torch.float32, device=dev())
    x16 = torch.empty((4096, 73), dtype=torch.float16, device=dev())
    xb16 = torch.empty((4096, 73), dtype=torch.bfloat16, device=dev())
    for buf, dt in ((x32, L.DT_F32), (x16, L.DT_F16), (xb16, L.DT_BF16)):
        L.check(h, L.load_library().nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, L.ptr(buf), dt, L.stream_ptr(dev())))
    assert torch.equal(x16, x32.to(torch.float16)) and torch.equal(xb16, x32.to(torch.bfloat16))


def test_gather_tile_kernel_equals_flat_kernel():
    """K1's TMA-staged tile kernel (rows of 128 texels, 2-D, step <= 1) is bit-identical to the flat kernel for fp32,
    f16 and bf16 X, for a full frame, for crops with device origins (training shape) and at steps 1/4, 1/2 and 1."""
    import ctypes as C
    n = nic()
    L = n._lib
    size = 512
    grids = I.make_grids(size, 2, seed=78)
    fp = [T(a) for a in grids]
    h, lib = L.handle(dev()), L.load_library()
    rng = np.random.default_rng(79)
    cases = [  # (fl, mip, block, num_blocks, origins)
        (0, 0, (512, 512), 1, None), (0, 0, (40, 384), 1, (9, 128)), (0, 1, (256, 256), 1, None), (0, 2, (128, 128), 1, None),
        (0, 0, (256, 256), 3, rng.integers(0, 257, (3, 2))), (0, 1, (128, 128), 5, rng.integers(0, 129, (5, 2))),
        (0, 2, (64, 128), 1, (64, 0)),
        # block heights that are not multiples of 8: the 16-bit kernel's 4-row super-tiles at steps 1/4 and 1/2
        (0, 0, (12, 256), 1, (3, 128)), (0, 1, (20, 128), 2, rng.integers(0, 100, (2, 2)))]
    for fl, mip, block, nb, org in cases:
        g0, g1 = fp[2 * fl], fp[2 * fl + 1]
        o0 = org if (org is not None and nb == 1 and not isinstance(org, np.ndarray)) else None
        geom = L.make_geom(L.METHOD_2D, g0, g1, block, nb, mip - 2 * (fl + 1), mip, 6, L.PE_TRIANGULAR, origin0=o0)
        coords = T(org, torch.int64) if isinstance(org, np.ndarray) else None
        N = nb * block[0] * block[1]
        for dt, td in ((L.DT_F32, torch.float32), (L.DT_F16, torch.float16), (L.DT_BF16, torch.bfloat16)):
            outs = []
            for flat in (0, 1):
                x = torch.full((N, 73), -7.0, dtype=td, device=dev())
                L.set_option(dev(), L.OPT_DISABLE_FAST2D, flat)
                try:
                    L.check(h, lib.nic_gather(h, C.byref(geom), L.ptr(g0), L.ptr(g1), L.ptr(coords), L.ptr(x), dt, L.stream_ptr(dev())))
                finally:
                    L.set_option(dev(), L.OPT_DISABLE_FAST2D, 0)
                outs.append(x)
            assert torch.equal(outs[0], outs[1]), (fl, mip, block, nb, td)
    # and against the oracle for the training shape
    coord = rng.integers(0, 257, (2, 2))
    X = n.image_compression.create_decoder_input_2d(fp, T(coord, torch.int64), 2, 0, 0).cpu().numpy()
    assert np.array_equal(X, O.create_decoder_input(grids, coord, 0, 0, 1))


def test_gather_full_frame_16bit_rows_are_the_fp32_rows_rounded_once():
    """K1 at BASELINE's full size (4096^2 texels, 2.45 GB of f16 / bf16 X): every value of the 16-bit tile kernel (patch
    staged already rounded, x-weighted G1 rows) equals the fp32 tile kernel's value rounded once — the size-independent form
    of the small-case oracle check above."""
    import ctypes as C
    n = nic()
    L = n._lib
    size = 4096
    grids = I.make_grids(size, 2, seed=80, no_mip=True, quantized=True)
    fp = [T(a) for a in grids]
    h, lib = L.handle(dev()), L.load_library()
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], size, 1, -2, 0, 6, L.PE_TRIANGULAR)
    x32 = torch.empty((size * size, 73), dtype=torch.float32, device=dev())
    L.check(h, lib.nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, L.ptr(x32), L.DT_F32, L.stream_ptr(dev())))
    for dt, td in ((L.DT_F16, torch.float16), (L.DT_BF16, torch.bfloat16)):
        x16 = torch.full((size * size, 73), -7.0, dtype=td, device=dev())
        L.check(h, lib.nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, L.ptr(x16), dt, L.stream_ptr(dev())))
        rows = 1 << 20
        for r0 in range(0, size * size, rows):
            assert torch.equal(x16[r0:r0 + rows], x32[r0:r0 + rows].to(td)), (td, r0)
        del x16


def test_gather_scatter_adjoint():
    """<gather(G), dX> == <G, scatter(dX)> on the grid columns (linearity / transpose property)."""
    ic = nic().image_compression
    size = 256
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    for mip in (0, 3):
        fl = O.create_pyramid_mip_levels(size, size // 4)[mip]
        fp = [T(a).requires_grad_(True) for a in I.make_grids(size, 2, seed=5)]
        coord = torch.tensor([[0, 0], [(size >> mip) - (256 >> mip), 0]], device=dev())
        X = ic.create_decoder_input_2d(fp, coord, 2, fl, mip)
        dX = torch.randn(X.shape, device=dev(), generator=torch.Generator(device=dev()).manual_seed(1))
        (X * dX).sum().backward()
        lhs = float((X.detach()[:, :60].double() * dX[:, :60].double()).sum())
        rhs = float((fp[2 * fl].detach().double() * fp[2 * fl].grad.double()).sum() +
                    (fp[2 * fl + 1].detach().double() * fp[2 * fl + 1].grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
        assert fp[2 * (1 - fl)].grad is None


# ------------------------------------------------------------------------------------------------ decode
def test_decode_2d_f32_golden():
    ic = nic().image_compression
    z = load("gather_decode_2d.npz")
    size = int(z["image_size"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=6)
    fp = [T(z[f"grid{i}"]) for i in range(4)]
    dec = make_decoder([z[f"param{i}"] for i in range(6)])
    for mip in range(7):
        ref = z[f"decode_mip{mip}"]
        fused = ic.decode_image(fp, dec, mip, pr=False).cpu().numpy()
        assert fused.shape == ref.shape
        np.testing.assert_allclose(fused, ref, rtol=1e-5, atol=1e-6)
        with torch.no_grad():                                  # the reference's two-call sequence
            two = dec(ic.finally_decode_input_2d(fp, size >> mip, mip)).reshape(ref.shape).cpu().numpy()
        np.testing.assert_allclose(two, ref, rtol=1e-5, atol=1e-6)
        u8 = ic.decode(fp, dec, mip, out_dtype=torch.uint8).cpu().numpy()
        ref8 = O.quantize_to_bit(ref, 8).astype(np.uint8)
        assert (u8 != ref8).mean() < 1e-3                      # only rounding ties at 1e-7 distance can differ
    zl = load("decode_2d_lowbits.npz")
    for b in (4, 2):
        gq = [T(a) for a in I.make_grids(size, 2, bits=b, seed=20 + b, quantized=True)]
        np.testing.assert_allclose(ic.decode_image(gq, dec, 0, pr=False).cpu().numpy(), zl[f"decode_bits{b}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("method", [3, 4])
def test_decode_3d_f32_golden(method):
    ic = nic().image_compression
    z = load(f"gather_decode_3d_m{method}.npz")
    size, cin = int(z["image_size"]), int(z["cin"])
    configure(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=3)
    fp = [T(a) for a in I.make_grids(size, 3, seed=31)]
    dec = make_decoder(I.make_mlp(cin, seed=32, gain=2.0))
    for mip in range(1, 6):
        out = ic.decode_image(fp, dec, mip, pr=False).cpu().numpy()
        np.testing.assert_allclose(out, z[f"decode_mip{mip}"], rtol=1e-5, atol=1e-6)
    out = ic.decode_image(fp, dec, 0, pr=False).cpu().numpy()
    np.testing.assert_allclose(out.reshape(-1)[z["decode_mip0_idx"]], z["decode_mip0_val"], rtol=1e-5, atol=1e-6)


def _tc_check(out_tc, out_ref, target255):
    """North-star tolerance of the tensor-core path against the fp32 reference output."""
    u_tc = np.floor(out_tc.astype(np.float64) * 255 + 0.5).astype(np.uint8) if out_tc.dtype != np.uint8 else out_tc
    u_ref = O.quantize_to_bit(out_ref, 8).astype(np.uint8)
    within1, same, worst = lsb_stats(u_tc, u_ref)
    dpsnr = abs(psnr256(u_tc, target255) - psnr256(u_ref, target255))
    return within1, same, worst, dpsnr


@pytest.mark.parametrize("prec", ["f16", "bf16"])
def test_decode_2d_tensor_core_golden(prec):
    ic = nic().image_compression
    z = load("gather_decode_2d.npz")
    size = int(z["image_size"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=6)
    fp = [T(z[f"grid{i}"]) for i in range(4)]
    dec = make_decoder([z[f"param{i}"] for i in range(6)])
    target = I.make_image(size, 2, seed=3).transpose(1, 2, 0) * 255
    for mip in range(7):
        ref = z[f"decode_mip{mip}"]
        out = ic.decode(fp, dec, mip, precision=prec).cpu().numpy()
        assert out.shape == ref.shape
        tol = 4e-3 if prec == "f16" else 2e-2
        assert np.abs(out - ref).max() < tol, (mip, np.abs(out - ref).max())
        s = size >> mip
        within1, same, worst, dpsnr = _tc_check(out, ref, target[:s, :s])
        if s >= 8:
            assert within1 >= 0.999 and dpsnr <= 0.05, (mip, within1, same, worst, dpsnr)
        u8 = ic.decode(fp, dec, mip, precision=prec, out_dtype=torch.uint8).cpu().numpy()
        assert np.array_equal(u8, np.floor(out.astype(np.float32) * np.float32(255) + np.float32(0.5)).astype(np.uint8))


@pytest.mark.parametrize("method", [3, 4])
def test_decode_3d_tensor_core(method):
    ic = nic().image_compression
    z = load(f"gather_decode_3d_m{method}.npz")
    size, cin = int(z["image_size"]), int(z["cin"])
    configure(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=3)
    fp = [T(a) for a in I.make_grids(size, 3, seed=31)]
    dec = make_decoder(I.make_mlp(cin, seed=32, gain=2.0))
    target = np.zeros((size, size, size, 3))
    for mip in range(1, 5):
        ref = z[f"decode_mip{mip}"]
        out = ic.decode(fp, dec, mip, precision="f16").cpu().numpy()
        s = size >> mip
        within1, same, worst, dpsnr = _tc_check(out, ref, target[:s, :s, :s])
        assert np.abs(out - ref).max() < 4e-3 and within1 >= 0.999, (mip, within1, same, worst)


def test_decode_vs_oracle_512_and_tiled_equals_single():
    """Config-1 shape (512^2): fp32 decode vs the oracle; tile-sharded decode equals the single-shot decode."""
    ic = nic().image_compression
    size = 512
    configure(IMAGE_SIZE=size)
    grids = I.make_grids(size, 2, seed=90, no_mip=True, quantized=True)
    params = I.make_mlp(73, seed=91, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    table = O.create_pyramid_mip_levels(size, size // 4)
    full = ic.decode(fp, dec, 0)
    ref = O.decode_block(grids, params, size, 0, table, 1)
    np.testing.assert_allclose(full.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    for prec in ("f32", "f16"):
        whole = ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8)
        parts = torch.empty_like(whole)
        for r0 in range(0, size, 128):                          # row tiles, as the multi-GPU decode shards them
            parts[r0:r0 + 128] = ic.decode(fp, dec, 0, size=(128, size), origin=(r0, 0), precision=prec, out_dtype=torch.uint8)
        assert torch.equal(parts, whole)
    tc = ic.decode(fp, dec, 0, precision="f16").cpu().numpy()
    within1, same, worst, dpsnr = _tc_check(tc, ref, I.make_image(size, 2, seed=3).transpose(1, 2, 0) * 255)
    assert within1 >= 0.999 and dpsnr <= 0.05, (within1, same, worst, dpsnr)


def test_fast2d_path_matches_general_kernel():
    """The aligned full-resolution 2-D fast path (aliased-cell MMA + selector matrix) against the general tensor-core
    kernel and the oracle, on blocks that are / are not eligible for it."""
    n = nic()
    ic, L = n.image_compression, n._lib
    size = 512
    configure(IMAGE_SIZE=size)
    grids = I.make_grids(size, 2, seed=94, no_mip=True, quantized=True)
    params = I.make_mlp(73, seed=95, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    table = O.create_pyramid_mip_levels(size, size // 4)
    ref = O.decode_block(grids, params, size, 0, table, 1)
    ref8 = O.quantize_to_bit(ref, 8).astype(np.uint8)
    for prec in ("f16", "bf16"):
        fast = ic.decode(fp, dec, 0, precision=prec).cpu().numpy()
        L.set_option(dev(), L.OPT_DISABLE_FAST2D, 1)
        try:
            gen = ic.decode(fp, dec, 0, precision=prec).cpu().numpy()
        finally:
            L.set_option(dev(), L.OPT_DISABLE_FAST2D, 0)
        tol = 4e-3 if prec == "f16" else 2.5e-2
        assert np.abs(fast - ref).max() < tol and np.abs(gen - ref).max() < tol, (np.abs(fast - ref).max(), np.abs(gen - ref).max())
        for o in (fast, gen):
            within1, same, worst = lsb_stats(np.floor(o.astype(np.float64) * 255 + 0.5).astype(np.uint8), ref8)
            assert within1 >= 0.999, (prec, within1, same, worst)
    # an eligible sub-block at a non-zero aligned origin, and a non-eligible (unaligned) one, agree with the full frame
    whole = ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8)
    # GELU on MUFU.TANH for every activation (NIC_OPT_GELU_POLY = 0, the round-1 kernel) against the default mix of
    # MUFU and polynomial activations: both within +-1 LSB of the oracle, and of each other
    L.set_option(dev(), L.OPT_GELU_POLY, 0)
    try:
        mufu = ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8)
    finally:
        L.set_option(dev(), L.OPT_GELU_POLY, -1)
    assert int((mufu.int() - whole.int()).abs().max()) <= 1
    within1, same, worst = lsb_stats(mufu.cpu().numpy(), ref8)
    assert within1 >= 0.999 and worst <= 1, (within1, same, worst)
    # frames with fewer tiles than slots / SMs, and tile counts that are not multiples of the slot count
    for sx, sy in ((8, 16), (8, 48), (24, 16), (40, 80), (448, 16)):
        part = ic.decode(fp, dec, 0, size=(sx, sy), origin=(64, 32), precision="f16", out_dtype=torch.uint8)
        assert torch.equal(part, whole[64:64 + sx, 32:32 + sy]), (sx, sy)
    a = ic.decode(fp, dec, 0, size=(64, 128), origin=(200, 304), precision="f16", out_dtype=torch.uint8)      # fast path
    assert torch.equal(a, whole[200:264, 304:432])
    b = ic.decode(fp, dec, 0, size=(64, 128), origin=(201, 300), precision="f16", out_dtype=torch.uint8)      # general path
    d = (b.int() - whole[201:265, 300:428].int()).abs()
    assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.97


def test_decode_4096_properties():
    """BASELINE config 2 shape (4096^2 full frame): properties that do not need the oracle at full size."""
    ic = nic().image_compression
    size = 4096
    configure(IMAGE_SIZE=size)
    grids = I.make_grids(size, 2, seed=92, no_mip=True, quantized=True)
    params = I.make_mlp(73, seed=93, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    whole = ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8)
    assert tuple(whole.shape) == (size, size, 3)
    # (a) any sub-block decoded on its own equals the same region of the full frame (tile independence)
    for (x0, y0, sx, sy) in ((0, 0, 128, 128), (4096 - 96, 4096 - 160, 96, 160), (1000, 2000, 333, 77)):
        part = ic.decode(fp, dec, 0, size=(sx, sy), origin=(x0, y0), precision="f16", out_dtype=torch.uint8)
        if x0 % 8 == 0 and y0 % 16 == 0 and sx % 8 == 0 and sy % 16 == 0:      # same (aligned fast-path) kernel: bit-equal
            assert torch.equal(part, whole[x0:x0 + sx, y0:y0 + sy])
        else:                                   # general kernel vs fast path: same image within the 1-LSB tolerance
            d = (part.int() - whole[x0:x0 + sx, y0:y0 + sy].int()).abs()
            assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.97
    # (b) a 256^2 window agrees with the oracle within the tensor-core tolerance
    table = O.create_pyramid_mip_levels(size, size // 4)
    ref = O.decode_block(grids, params, 256, 0, table, 1, origin=(1792, 3840))
    win = whole[1792:2048, 3840:4096].cpu().numpy()
    within1, same, worst = lsb_stats(win, O.quantize_to_bit(ref, 8).astype(np.uint8))
    assert within1 >= 0.999, (within1, same, worst)
    # (c) the fp32 path on the same window meets 1e-5
    f32 = ic.decode(fp, dec, 0, size=256, origin=(1792, 3840)).cpu().numpy()
    np.testing.assert_allclose(f32, ref, rtol=1e-5, atol=1e-6)


def test_decode_bands_tile_the_frame():
    """Multi-GPU decode partitioning (parallel.decode_band) emulated on one device: the bands of 1, 2, 3 and 8 ranks
    are bit-equal to the rows of the single-GPU frame (no collective, no halo)."""
    n = nic()
    ic, par = n.image_compression, n.parallel
    size = 512
    configure(IMAGE_SIZE=size)
    fp = [T(a) for a in I.make_grids(size, 2, seed=70, no_mip=True, quantized=True)]
    dec = make_decoder(I.make_mlp(73, seed=71, gain=2.0))
    whole = ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8)
    for world in (1, 2, 3, 8):
        rows_seen = 0
        for rank in range(world):
            row0, band = par.decode_band(fp, dec, 0, rank=rank, world=world, precision="f16", out_dtype=torch.uint8)
            assert torch.equal(band, whole[row0:row0 + band.shape[0]])
            rows_seen += band.shape[0]
        assert rows_seen == size


def test_random_access_decode_3d():
    """BASELINE config 4 shape (3-D LUT, random-access queries): each query is a 1x1x1 block whose origin is the
    query coordinate; results equal the dense decode at those coordinates."""
    n = nic()
    ic, L = n.image_compression, n._lib
    import ctypes as C
    size = 64
    configure(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=3, CROP_MIP_LEVEL=5)
    grids = I.make_grids(size, 3, seed=72, no_mip=True, quantized=True)
    params = I.make_mlp(127, seed=73, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    dense32 = ic.decode(fp, dec, 0, precision="f32")
    dense16 = ic.decode(fp, dec, 0, precision="f16")
    rng = np.random.default_rng(74)
    q = rng.integers(0, size, (10007, 3))
    for prec, dense, tol in (("f32", dense32, 0.0), ("f16", dense16, 0.0)):
        out = ic.decode_points(fp, dec, torch.tensor(q), 0, precision=prec)
        want = dense[q[:, 0], q[:, 1], q[:, 2]]
        assert float((out - want).abs().max()) <= tol, prec
    assert ic.decode_points(fp, dec, torch.zeros((0, 3), dtype=torch.int64), 0).shape == (0, 3)


# ------------------------------------------------------------------------------------------------ training
def test_two_call_autograd_matches_golden_first_step():
    """create_decoder_input_2d -> + noise -> decoder -> MSE -> backward, as train_models does, with torch autograd."""
    ic = nic().image_compression
    z = load("train_2d.npz")
    size, nc = int(z["size"]), int(z["nc"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=9)
    lod = int(z["lods"][0])
    fp = [T(a).requires_grad_(True) for a in I.make_grids(size, 2, seed=40)]
    dec = make_decoder(I.make_mlp(73, seed=41))
    mips = I.box_mips(I.make_image(size, 2, seed=42), 9)
    crop = 2 ** (8 - lod)
    coord = z["coord0"]
    target = np.concatenate([mips[lod][:, c[0]:c[0] + crop, c[1]:c[1] + crop].reshape(3, -1).T for c in coord], 0)
    fl = ic.feature_pyramid_mip_levels()[lod]
    x = ic.create_decoder_input_2d(fp, T(coord), nc, fl, lod)
    out = dec(x + T(I.make_noise(nc * crop * crop, 73, 8, 1000)))
    loss = torch.nn.functional.mse_loss(out, T(target))
    loss.backward()
    assert abs(float(loss) - float(z["losses"][0])) <= 2e-5 * float(z["losses"][0])
    np.testing.assert_allclose(out.detach().cpu().numpy().reshape(-1)[z["out0_idx"]], z["out0_val"], rtol=1e-5, atol=1e-6)
    for i, p in enumerate(dec.parameters_list()):
        np.testing.assert_allclose(p.grad.cpu().numpy(), z[f"grad0_param{i}"], rtol=2e-3, atol=2e-7)
    np.testing.assert_allclose(fp[2 * fl].grad.cpu().numpy().reshape(-1)[z["grad0_g0_idx"]], z["grad0_g0_val"], rtol=2e-3, atol=1e-8)
    np.testing.assert_allclose(fp[2 * fl + 1].grad.cpu().numpy().reshape(-1)[z["grad0_g1_idx"]], z["grad0_g1_val"], rtol=2e-3, atol=1e-8)


def _run_fused(z, grids, params, mips, cin, method, dim, crop_level, noise_seed0, table=None):
    ic = nic().image_compression
    nc, t_max = int(z["nc"]), int(z["t_max"])
    lods = [int(v) for v in z["lods"]]
    fp = [T(a) for a in grids]
    dec = make_decoder(params)
    tr = ic.FusedTrainer(fp, dec, num_epochs=t_max, fp_bits=8)
    losses = []
    for e, lod in enumerate(lods):
        crop = 2 ** max(0, (8 if dim == 2 else crop_level) - lod)
        coord = z[f"coord{e}"]
        tg = []
        for c in coord:
            sl = (slice(None),) + tuple(slice(int(c[a]), int(c[a]) + crop) for a in range(dim))
            tg.append(mips[lod][sl].reshape(3, -1).T)
        noise = T(I.make_noise(nc * crop ** dim, cin, 8, noise_seed0 + e))
        losses.append(float(tr.step(T(coord), T(np.concatenate(tg, 0)), lod, noise=noise)))
    return losses, [g.cpu().numpy() for g in tr.fp], [p.detach().cpu().numpy() for p in dec.parameters_list()], tr


def test_fused_training_2d_sequence_golden():
    """40 fused steps over mixed LODs, including the freeze + quantise switch at epoch > 0.95 N."""
    z = load("train_2d.npz")
    size = int(z["size"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=9)
    grids = I.make_grids(size, 2, seed=40)
    losses, fp, pr, tr = _run_fused(z, grids, I.make_mlp(73, seed=41), I.box_mips(I.make_image(size, 2, seed=42), 9), 73, 1, 2, 8, 1000)
    np.testing.assert_allclose(losses, z["losses"], rtol=2e-4)
    for i in range(6):
        np.testing.assert_allclose(pr[i], z[f"param{i}"], rtol=0, atol=2e-3)
    for i in range(len(fp)):
        # the grids were quantised at the freeze: a value that sits on a rounding boundary may land one code (1/255) away
        # when the float atomics of the scatter add up in another order — allowed for at most 0.1 % of the sampled values
        d = np.abs(fp[i].reshape(-1)[z[f"grid{i}_idx"]] - z[f"grid{i}_val"])
        assert d.max() <= 1.0 / 255 + 2e-3 and (d <= 2e-3).mean() >= 0.999, (i, d.max(), (d <= 2e-3).mean())
    for i in (6, 7):       # never-active level: untouched by Adam, then quantised -> bit-exact
        assert np.array_equal(fp[i].reshape(-1)[z[f"grid{i}_idx"]], z[f"grid{i}_val"])
    # per-tensor step counts (reference: grids of inactive levels keep t = 0)
    counts = {k: v[2] for k, v in tr.state.items()}
    assert counts[("p", 0)] == 40 and ("g", 6) not in counts and counts[("g", 0)] < 40


def test_fused_training_default_shape_golden():
    z = load("train_2d_lod0.npz")
    size = int(z["size"])
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=9)
    grids = I.make_grids(size, 2, seed=40)
    losses, fp, pr, _ = _run_fused(z, grids, I.make_mlp(73, seed=41), I.box_mips(I.make_image(size, 2, seed=42), 9), 73, 1, 2, 8, 2000)
    np.testing.assert_allclose(losses, z["losses"], rtol=2e-4)
    for i in range(6):
        np.testing.assert_allclose(pr[i], z[f"param{i}"], rtol=0, atol=2e-3)
    for i in (0, 1):
        np.testing.assert_allclose(fp[i].reshape(-1)[z[f"grid{i}_idx"]], z[f"grid{i}_val"], rtol=0, atol=2e-3)


@pytest.mark.parametrize("method", [3, 4])
def test_fused_training_3d_golden(method):
    z = load(f"train_3d_m{method}.npz")
    size, cin = int(z["size"]), int(z["cin"])
    configure(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=3)
    grids = I.make_grids(size, 3, seed=50)
    losses, fp, pr, _ = _run_fused(z, grids, I.make_mlp(cin, seed=51), [z["target"]] * 6, cin, method, 3, 3, 3000)
    np.testing.assert_allclose(losses, z["losses"], rtol=2e-4)
    for i in range(6):
        np.testing.assert_allclose(pr[i], z[f"param{i}"], rtol=0, atol=2e-3)
    for i in range(len(fp)):
        np.testing.assert_allclose(fp[i], z[f"grid{i}"], rtol=0, atol=2e-3)


def test_fused_step_gradients_vs_oracle_fp64():
    """One fused step against the oracle's fp64 hand-derived backward (tight tolerance on the first update)."""
    n = nic()
    ic, L = n.image_compression, n._lib
    import ctypes as C
    size = 256
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    grids = I.make_grids(size, 2, seed=60)
    params = I.make_mlp(73, seed=61, gain=1.5)
    mip, fl, nc, crop = 2, 0, 3, 64
    coord = np.array([[0, 0], [0, 0], [0, 0]])
    img = I.box_mips(I.make_image(size, 2, seed=62), 8)[mip]
    target = np.concatenate([img[:, :crop, :crop].reshape(3, -1).T] * nc, 0)
    noise = I.make_noise(nc * crop * crop, 73, 8, 63)
    loss, out, grads, dg0, dg1 = O.train_forward_backward(grids, params, coord, target, fl, mip, 1, noise)
    fp = [T(a) for a in grids]
    pt = [T(p) for p in params]
    m = L.make_mlp(pt)
    g = [torch.zeros_like(p) for p in pt]
    gm = L.make_mlp_grad(g)
    d0, d1 = torch.zeros_like(fp[0]), torch.zeros_like(fp[1])
    ls = torch.zeros(4, device=dev())
    o = torch.empty((nc * crop * crop, 3), device=dev())
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], crop, nc, 0, mip, 6, L.PE_TRIANGULAR)
    h = L.handle(dev())
    coord_t, target_t, noise_t = T(coord), T(target), T(noise)      # keep the device buffers alive across the launch
    L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), L.ptr(coord_t), C.byref(m),
                                               L.ptr(target_t), L.ptr(noise_t), 0, 0, 0, 0, C.byref(gm), L.ptr(d0), L.ptr(d1),
                                               L.ptr(ls), L.ptr(o), L.PREC_F32, L.stream_ptr(dev())))
    n_all = nc * crop * crop * 3
    assert abs(float(ls[0]) / n_all - loss) <= 1e-5 * loss
    np.testing.assert_allclose(o.cpu().numpy(), out, rtol=1e-5, atol=1e-6)
    for t, k in zip(g, ("W1", "b1", "W2", "b2", "W3", "b3")):
        np.testing.assert_allclose(t.cpu().numpy(), grads[k], rtol=1e-3, atol=1e-8)
    np.testing.assert_allclose(d0.cpu().numpy(), dg0, rtol=1e-3, atol=1e-9)
    np.testing.assert_allclose(d1.cpu().numpy(), dg1, rtol=1e-3, atol=1e-9)


def test_philox_noise_distribution_and_determinism():
    """In-kernel noise: (U[0,1) - .5) / 2^bits, deterministic in (seed, step).  Checked on the per-sample decoder
    outputs of the first step (per-sample arithmetic is deterministic; the loss sum is reduced with float atomics)."""
    ic = nic().image_compression
    size = 256
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)

    def run(seed, noise):
        fp = [T(a) for a in I.make_grids(size, 2, seed=60)]
        dec = make_decoder(I.make_mlp(73, seed=61, gain=2.0))
        tr = ic.FusedTrainer(fp, dec, num_epochs=100, fp_bits=4, seed=seed)
        img = I.box_mips(I.make_image(size, 2, seed=62), 8)[2]
        tg = T(img[:, :64, :64].reshape(3, -1).T)
        out = torch.empty((64 * 64, 3), dtype=torch.float32, device=dev())
        loss = float(tr.step(torch.tensor([[0, 0]]), tg, 2, noise=noise, out=out))
        return loss, out.cpu().numpy()

    (la, a), (lb, b), (lc, c), (l0, o0) = run(1, None), run(1, None), run(2, None), run(1, False)
    assert np.array_equal(a, b) and abs(la - lb) <= 1e-5 * la       # same (seed, step): same noise
    assert not np.array_equal(a, c)                                  # another seed: another draw
    da, dc = a - o0, c - o0
    assert 1e-5 < np.abs(da).mean() < 2e-2                           # noise of +-2^-5 on 73 inputs: small, non-zero
    assert abs(da.mean()) < 0.05 * np.abs(da).mean() + 1e-6          # zero-mean perturbation
    assert abs(np.corrcoef(da.reshape(-1), dc.reshape(-1))[0, 1]) < 0.1   # independent streams per seed
    assert abs(la - l0) < 0.05 * l0


# ------------------------------------------------------------------------------------------------ tensor-core training step
def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("prec", ["f16", "bf16"])
@pytest.mark.parametrize("case", ["mip2_aligned", "mip0_unaligned", "mip5_small_ragged", "large_grid", "mip2_cout5"])
def test_train_tc_step_vs_oracle_fp64(prec, case):
    """The tcgen05 training step against the oracle's fp64 backward.  Tolerances are those of 16-bit operands with fp32
    accumulation (tanh-form GELU in forward and backward): loss 1e-2 relative; every gradient tensor within a few % in
    relative L2 norm.  Noise is injected (the same tensor on both sides)."""
    n = nic()
    L = n._lib
    import ctypes as C
    size = 4096 if case == "large_grid" else 256       # 1025^2 nodes: the tiled (bandwidth-shaped) relayout kernels
    grids = I.make_grids(size, 2, seed=60, no_mip=case == "large_grid")
    cout = 5 if case == "mip2_cout5" else 3
    params = I.make_mlp(73, cout=cout, seed=61, gain=1.5)
    rng = np.random.default_rng(64)
    if case == "large_grid":
        mip, fl, nc, crop = 0, 0, 2, 128
        coord = rng.integers(0, size - crop + 1, (nc, 2))
    elif case in ("mip2_aligned", "mip2_cout5"):
        mip, fl, nc, crop = 2, 0, 3, 64
        coord = np.array([[0, 0], [0, 0], [0, 0]])
    elif case == "mip0_unaligned":
        mip, fl, nc, crop = 0, 0, 2, 128            # step 1/4: 4 (G0) and 8 (G1) lanes share a node, origins unaligned
        coord = rng.integers(0, size - crop + 1, (nc, 2))
    else:
        mip, fl, nc, crop = 5, 1, 5, 5              # step 2 (no interpolation), N = 125: one ragged tile
        coord = rng.integers(0, (size >> mip) - crop + 1, (nc, 2))
    img = I.box_mips(I.make_image(512 if case == "large_grid" else size, 2, seed=62), 8)[mip]
    target = np.concatenate([img[:, c[0] % 256:c[0] % 256 + crop, c[1] % 256:c[1] % 256 + crop].reshape(3, -1).T for c in coord], 0)
    if cout != 3:       # multi-channel material stack: extra target channels
        target = np.concatenate([target, 1.0 - target], 1)[:, :cout].astype(np.float32)
    noise = I.make_noise(nc * crop * crop, 73, 8, 63)
    loss, out, grads, dg0, dg1 = O.train_forward_backward(grids, params, coord, target, fl, mip, 1, noise, size=crop,
                                                          dtype=np.float32 if case == "large_grid" else np.float64)
    fp = [T(a) for a in grids]
    pt = [T(p) for p in params]
    m = L.make_mlp(pt)
    g = [torch.zeros_like(p) for p in pt]
    gm = L.make_mlp_grad(g)
    g0t, g1t = fp[2 * fl], fp[2 * fl + 1]
    d0, d1 = torch.zeros_like(g0t), torch.zeros_like(g1t)
    ls = torch.zeros(4, device=dev())
    o = torch.empty((nc * crop * crop, cout), device=dev())
    geom = L.make_geom(L.METHOD_2D, g0t, g1t, crop, nc, mip - 2 * (fl + 1), mip, 6, L.PE_TRIANGULAR)
    h = L.handle(dev())
    coord_t, target_t, noise_t = T(coord, torch.int64), T(target), T(noise)
    for rep in range(2):            # twice: the second call checks that the private gradient scratch was left zeroed
        for t in g + [d0, d1, ls]:
            t.zero_()
        L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(g0t), L.ptr(g1t), L.ptr(coord_t), C.byref(m),
                                                   L.ptr(target_t), L.ptr(noise_t), 0, 0, 0, 0, C.byref(gm), L.ptr(d0),
                                                   L.ptr(d1), L.ptr(ls), L.ptr(o), L.PRECISIONS[prec], L.stream_ptr(dev())))
        n_all = nc * crop * crop * cout
        tol = 1.0 if prec == "f16" else 4.0
        assert abs(float(ls[0]) / n_all - loss) <= 1e-2 * tol * loss
        assert np.abs(o.cpu().numpy() - out).max() <= 4e-3 * tol
        for t, k in zip(g, ("W1", "b1", "W2", "b2", "W3", "b3")):
            assert _rel_l2(t.cpu().numpy(), grads[k]) <= 2e-2 * tol, (k, _rel_l2(t.cpu().numpy(), grads[k]))
        assert _rel_l2(d0.cpu().numpy(), dg0) <= 2e-2 * tol, _rel_l2(d0.cpu().numpy(), dg0)
        assert _rel_l2(d1.cpu().numpy(), dg1) <= 2e-2 * tol, _rel_l2(d1.cpu().numpy(), dg1)
        # nodes outside the crop footprints receive exactly nothing
        assert np.array_equal(d0.cpu().numpy() == 0, dg0 == 0) or np.all((d0.cpu().numpy() != 0) <= (dg0 != 0))


def test_train_tc_mlp_gradients_deterministic_and_phase_counters():
    """The tensor-core step keeps the decoder gradients as per-CTA sums in tensor memory and adds the CTAs' slices in a
    fixed order; the three tile slots of a CTA accumulate into the same accumulators in the order the tensor pipe takes
    their batches, so two runs on the same inputs agree to fp32 rounding (static or dynamic tile order), while outputs and
    per-sample results are bit-identical.  (The grid gradients are scattered with float atomics: rounding, too.)  Also
    exercises nic_debug_counters: the phase profile counts every tile exactly once."""
    n = nic()
    L = n._lib
    import ctypes as C
    size, mip, fl, nc, crop = 256, 0, 0, 6, 128
    grids = I.make_grids(size, 2, seed=90)
    pt = [T(p) for p in I.make_mlp(73, seed=91, gain=1.5)]
    fp = [T(a) for a in grids]
    rng = np.random.default_rng(92)
    coord_t = T(rng.integers(0, size - crop + 1, (nc, 2)), torch.int64)
    target_t = T(rng.random((nc * crop * crop, 3), dtype=np.float32))
    noise_t = T(I.make_noise(nc * crop * crop, 73, 8, 93))
    g0t, g1t = fp[0], fp[1]
    geom = L.make_geom(L.METHOD_2D, g0t, g1t, crop, nc, mip - 2 * (fl + 1), mip, 6, L.PE_TRIANGULAR)
    m = L.make_mlp(pt)
    h = L.handle(dev())
    runs = []
    L.set_option(dev(), L.OPT_DEBUG_KNOCKOUT, 8)
    L.debug_counters(dev())
    try:
        for rep in range(3):
            L.set_option(dev(), L.OPT_STATIC_TILES, 1 if rep < 2 else 0)
            g = [torch.zeros_like(p) for p in pt]
            gm = L.make_mlp_grad(g)
            d0, d1 = torch.zeros_like(g0t), torch.zeros_like(g1t)
            ls = torch.zeros(4, device=dev())
            o = torch.empty((nc * crop * crop, 3), device=dev())
            L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(g0t), L.ptr(g1t), L.ptr(coord_t), C.byref(m),
                                                       L.ptr(target_t), L.ptr(noise_t), 0, 0, 0, 0, C.byref(gm), L.ptr(d0),
                                                       L.ptr(d1), L.ptr(ls), L.ptr(o), L.PREC_F16, L.stream_ptr(dev())))
            counters = L.debug_counters(dev())
            assert counters[15] == nc * crop * crop // 128               # every tile of 128 samples, once
            assert all(c > 0 for c in counters[:13])
            runs.append(([t.cpu().numpy() for t in g], o.cpu().numpy(), d0.cpu().numpy(), d1.cpu().numpy()))
    finally:
        L.set_option(dev(), L.OPT_DEBUG_KNOCKOUT, 0)
        L.set_option(dev(), L.OPT_STATIC_TILES, 0)
    for a, b, c in zip(runs[0][0], runs[1][0], runs[2][0]):
        assert _rel_l2(b, a) < 1e-5 and _rel_l2(c, a) < 1e-5         # the same sums in another order
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][1], runs[2][1])
    assert _rel_l2(runs[0][2], runs[1][2]) < 1e-5 and _rel_l2(runs[0][3], runs[1][3]) < 1e-5


def test_train_tc_short_run_tracks_f32_path():
    """200 fused steps on a 512^2 synthetic image, once on the fp32 path and once on the f16 tensor-core path with the
    same crops and LODs: the final full-frame PSNR (reference formula) agrees within 0.05 dB (north-star tolerance)."""
    n = nic()
    ic = n.image_compression
    size, steps = 512, 200
    img = I.make_image(size, 2, seed=80)
    target8 = torch.tensor(np.floor(img * 255 + 0.5).astype(np.uint8)).permute(1, 2, 0).contiguous().to(dev())
    rng = np.random.default_rng(81)
    coords = rng.integers(0, size - 256 + 1, (steps, 8, 2))
    img_t = T(img)
    psnr = {}
    for prec in ("f32", "f16", "bf16"):
        configure(IMAGE_SIZE=size)
        grids = I.make_grids(size, 2, seed=82, no_mip=True)
        fp = [T(a) for a in grids]
        dec = make_decoder(I.make_mlp(73, seed=83))
        tr = ic.FusedTrainer(fp, dec, num_epochs=steps, fp_bits=8, seed=5, precision=prec)
        for s in range(steps):
            c = coords[s]
            tg = torch.stack([img_t[:, a:a + 256, b:b + 256].reshape(3, -1).T for a, b in c])
            tr.step(torch.tensor(c), tg, 0)
        out8 = ic.decode(tr.fp, dec, 0, precision="f32", out_dtype=torch.uint8)
        psnr[prec] = n.utils.calculate_psnr(target8, out8)
    assert psnr["f32"] > 18.0, psnr
    assert abs(psnr["f16"] - psnr["f32"]) <= 0.05, psnr
    assert abs(psnr["bf16"] - psnr["f32"]) <= 0.05, psnr         # bf16 operands: the same north-star tolerance


def test_random_crop_dataset_targets():
    """random_crop_dataset (image_compression.py:26-50): the one-kernel target gather equals the reference's per-crop
    slice / reshape / transpose, in 2-D and 3-D, and the LOD / origin draws stay inside the image."""
    import random
    ic = nic().image_compression
    configure(IMAGE_SIZE=64, TF_NO_MIP=False, MAX_MIP_LEVEL=6)
    rng = np.random.default_rng(90)
    for dim in (2, 3):
        size = 64 if dim == 2 else 16
        imgs = [T(rng.random((3,) + (size >> m,) * dim).astype(np.float32)) for m in range(4)]
        configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=3)
        random.seed(3)
        torch.manual_seed(4)
        for _ in range(6):
            tg, coord, lod = ic.random_crop_dataset(imgs, 8, 5, False, dim=dim)
            s = max(1, 8 // 2 ** lod)
            assert tuple(tg.shape) == (5, s ** dim, 3) and 0 <= lod <= 3
            for k, c in enumerate(coord.tolist()):
                sl = (slice(None),) + tuple(slice(c[a], c[a] + s) for a in range(dim))
                assert torch.equal(tg[k], imgs[lod][sl].reshape(3, -1).T)


@pytest.mark.parametrize("bits", [8, 4, 2])
def test_sub_byte_code_packing_round_trip(bits):
    """fp_savable_packed / fp_load_packed: 8/bits codes per byte, lossless w.r.t. the reference's one-code-per-byte
    fp_savable (bit-exact codes), ragged sizes included; the reloaded grids equal quantize4fp of the originals."""
    fpd, models = nic().fp_def, nic().models
    rng = np.random.default_rng(91)
    lo, hi = I.q_range(bits)
    grids = [T(rng.uniform(lo, hi, shape).astype(np.float32)) for shape in ((12, 33, 33), (12, 17, 17), (3, 5, 7), (1, 1, 3))]
    packed = fpd.fp_savable_packed(grids, bits)
    ref_codes = fpd.fp_savable(grids, bits)
    for (p, shape), g, c in zip(packed, grids, ref_codes):
        assert p.numel() == (g.numel() * bits + 7) // 8 and shape == tuple(g.shape)
        per = 8 // bits
        want = np.zeros(p.numel() * per, np.uint8)
        want[:g.numel()] = c.cpu().numpy().reshape(-1)
        want = (want.reshape(-1, per).astype(np.uint32) << (np.arange(per) * bits)).sum(1).astype(np.uint8)
        assert np.array_equal(p.cpu().numpy(), want)
    for g, r in zip(grids, fpd.fp_load_packed(packed, bits)):
        assert torch.equal(r, models.quantize4fp(g, bits))


@pytest.mark.parametrize("cout", [1, 5, 9, 16])
def test_decode_multi_channel_outputs(cout):
    """BASELINE config 3 shape (multi-channel material stack): decoders with Cout != 3 on every decode path (fast 2-D
    tensor-core kernel, general tensor-core kernel, fp32 kernel) against the oracle."""
    ic = nic().image_compression
    size = 128
    configure(IMAGE_SIZE=size, OUTPUT_CHANNELS=cout)
    grids = I.make_grids(size, 2, seed=96, no_mip=True, quantized=True)
    params = I.make_mlp(73, cout=cout, seed=97, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    table = O.create_pyramid_mip_levels(size, size // 4)
    ref = O.decode_block(grids, params, size, 0, table, 1)
    assert ref.shape == (size, size, cout)
    f32 = ic.decode(fp, dec, 0, precision="f32").cpu().numpy()
    np.testing.assert_allclose(f32, ref, rtol=1e-5, atol=1e-6)
    ref8 = O.quantize_to_bit(ref, 8).astype(np.uint8)
    fast = ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8).cpu().numpy()                       # fast path
    gen = ic.decode(fp, dec, 0, size=(120, 120), origin=(3, 5), precision="f16", out_dtype=torch.uint8).cpu().numpy()
    for got, want in ((fast, ref8), (gen, ref8[3:123, 5:125])):
        within1, same, worst = lsb_stats(got, want)
        assert got.shape == want.shape and within1 >= 0.999, (cout, within1, same, worst)


@pytest.mark.parametrize("tc_prec", ["f16", "bf16"])
def test_train_tc_multi_lod_trajectory_tracks_f32(tc_prec):
    """FusedTrainer with mips on (alternating LODs, per-tensor Adam step counts, freeze + quantise at 95 %): the f16 /
    bf16 tensor-core trainer follows the fp32 trainer step for step — same crops, same injected noise, losses within
    2 % (f16) / 3 % (bf16: 8 mantissa bits in the operands), and identical Adam step counters (tensors of inactive
    levels are never touched)."""
    n = nic()
    ic = n.image_compression
    size, steps = 256, 24
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    mips = [T(m) for m in I.box_mips(I.make_image(size, 2, seed=84), 8)]
    lods = [0, 1, 2, 4, 0, 3, 5, 1, 6, 0, 2, 7, 0, 1, 4, 0, 2, 0, 1, 0, 3, 0, 1, 0]
    rng = np.random.default_rng(85)
    runs = {}
    for prec in ("f32", tc_prec):
        fp = [T(a) for a in I.make_grids(size, 2, seed=86)]
        dec = make_decoder(I.make_mlp(73, seed=87))
        tr = ic.FusedTrainer(fp, dec, num_epochs=steps, fp_bits=8, seed=9, precision=prec)
        r2 = np.random.default_rng(88)
        losses = []
        for s, lod in enumerate(lods):
            crop = 2 ** (8 - lod)
            dsize = size >> lod
            coord = r2.integers(0, dsize - crop + 1, (4, 2))
            tg = ic.sample_crops(mips[lod], torch.tensor(coord), crop)
            noise = T(I.make_noise(4 * crop * crop, 73, 8, 4000 + s)) if s < steps * 0.95 else None
            losses.append(float(tr.step(torch.tensor(coord), tg, lod, noise=noise)))
        runs[prec] = (losses, {k: v[2] for k, v in tr.state.items()}, [g.clone() for g in tr.fp], tr.frozen)
    l32, l16 = np.array(runs["f32"][0]), np.array(runs[tc_prec][0])
    tol = 0.02 if tc_prec == "f16" else 0.03
    assert np.all(np.abs(l16 - l32) <= tol * l32 + 1e-5), (l32, l16, np.max(np.abs(l16 - l32) / l32))
    assert runs["f32"][1] == runs[tc_prec][1]                      # per-tensor Adam step counts
    assert runs["f32"][3] and runs[tc_prec][3]                      # both froze + quantised the grids at 95 %
    for a, b in zip(runs["f32"][2], runs[tc_prec][2]):
        # quantised grids: codes may differ by one level where the two trajectories straddle a rounding boundary.  With
        # bf16 operands (8 mantissa bits) the sign of a near-zero gradient can flip, and ONE Adam step moves a grid value by
        # up to lr = 0.01 = 2.5 codes either way: >= 99 % of the codes within one level, none further than 8
        d = (a - b).abs()
        if tc_prec == "f16":
            assert float(d.max()) <= 1.0 / 255 + 1e-6
        else:
            assert float((d <= 1.0 / 255 + 1e-6).float().mean()) >= 0.99 and float(d.max()) <= 8.0 / 255 + 1e-6
        assert float(((a - b).abs() > 1e-6).float().mean()) < 0.05


def test_non_cubic_volume_and_slab_sharding():
    """BASELINE config 5 shape in miniature: a non-cubic volume (x, y, t) = (24, 16, 40) on non-cubic grids, decoded
    whole and as per-rank frame slabs (parallel.decode_slab, no collective); every 8^3 sub-cube equals the oracle's
    cube decode at that origin (the reference itself only decodes cubes), slabs are bit-equal to the whole volume."""
    n = nic()
    ic, par = n.image_compression, n.parallel
    configure(IMAGE_SIZE=64, IMAGE_DIMENSION=3, COMPRESSION_METHOD=3, CROP_MIP_LEVEL=3)
    vol = (24, 16, 40)
    rng = np.random.default_rng(92)
    lo, hi = I.q_range(8)
    # grids [C, z, y, x] with nodes = texels/4 + 1 (G0) and texels/8 + 1 (G1) per axis; x is the FIRST image axis
    g0 = rng.uniform(lo, hi, (12, vol[2] // 4 + 1, vol[1] // 4 + 1, vol[0] // 4 + 1)).astype(np.float32)
    g1 = rng.uniform(lo, hi, (12, vol[2] // 8 + 1, vol[1] // 8 + 1, vol[0] // 8 + 1)).astype(np.float32)
    grids = [g0, g1]
    params = I.make_mlp(127, seed=93, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    table = {0: 0}
    for prec, tol in (("f32", 2e-6), ("f16", 4e-3)):
        whole = ic.decode(fp, dec, 0, size=vol, precision=prec, level_table=table)
        assert tuple(whole.shape) == vol + (3,)
        for origin in ((0, 0, 0), (16, 8, 32), (8, 0, 16)):
            ref = O.decode_block(grids, params, 8, 0, table, 3, origin=origin)
            got = whole[origin[0]:origin[0] + 8, origin[1]:origin[1] + 8, origin[2]:origin[2] + 8].cpu().numpy()
            assert np.abs(got - ref).max() <= tol, (prec, origin, np.abs(got - ref).max())
        for world in (1, 3, 8):
            seen = 0
            for rank in range(world):
                f0, slab = par.decode_slab(fp, dec, vol, 0, rank=rank, world=world, precision=prec, level_table=table)
                assert torch.equal(slab, whole[f0:f0 + slab.shape[0]])
                seen += slab.shape[0]
            assert seen == vol[0]


def test_lut_65_cubed_extension():
    """BASELINE config 4 shape: a 65^3 LUT.  The reference cannot hold it (FEATURE_PYRAMID_SIZE = 65 // 4 = 16 gives 17
    nodes, texel 64 needs node 17: SURVEY 8(d)); with ONE more node per grid axis the same arithmetic covers it.  Dense
    decode of all 65^3 texels and random-access queries agree with the oracle evaluated on the same 18-/10-node grids."""
    n = nic()
    ic = n.image_compression
    configure(IMAGE_SIZE=64, IMAGE_DIMENSION=3, COMPRESSION_METHOD=3, CROP_MIP_LEVEL=3)
    rng = np.random.default_rng(94)
    lo, hi = I.q_range(8)
    grids = [rng.uniform(lo, hi, (12, 18, 18, 18)).astype(np.float32), rng.uniform(lo, hi, (12, 10, 10, 10)).astype(np.float32)]
    params = I.make_mlp(127, seed=95, gain=2.0)
    fp, dec = [T(a) for a in grids], make_decoder(params)
    table = {0: 0}
    ref = O.decode_block(grids, params, 65, 0, table, 3)
    dense = ic.decode(fp, dec, 0, size=65, precision="f32", level_table=table).cpu().numpy()
    np.testing.assert_allclose(dense, ref, rtol=1e-5, atol=1e-6)
    q = rng.integers(0, 65, (4099, 3))
    q[:3] = [[64, 64, 64], [0, 0, 64], [64, 0, 0]]
    pts = ic.decode_points(fp, dec, torch.tensor(q), 0, precision="f16", out_dtype=torch.uint8, level_table=table).cpu().numpy()
    want = O.quantize_to_bit(ref, 8).astype(np.uint8)[q[:, 0], q[:, 1], q[:, 2]]
    within1, same, worst = lsb_stats(pts, want)
    assert within1 >= 0.999, (within1, same, worst)


@pytest.mark.parametrize("prec", ["f16", "bf16"])
@pytest.mark.parametrize("mip", [0, 2])
@pytest.mark.parametrize("method", [3, 4])
def test_train_tc_step_3d_vs_oracle_fp64(prec, mip, method):
    """The tcgen05 training step on the 3-D methods against the oracle's fp64 backward, at step 1/4 (mip 0) and step 1
    (mip 2): method 3 (8 raw G0 corners, Cin 127 -> K = 128, one CTA per SM, triangular PE) and the "proposed" method 4
    (tetrahedral G0 corners, sinusoidal PE, Cin 79); both use the 8-corner AS-CODED G1 weights."""
    n = nic()
    L = n._lib
    import ctypes as C
    size = 64
    grids = I.make_grids(size, 3, seed=50)
    cin = 127 if method == 3 else 79
    params = I.make_mlp(cin, seed=51, gain=1.5)
    rng = np.random.default_rng(65)
    fl, nc, crop = 0, 3, 8
    coord = rng.integers(0, (size >> mip) - crop + 1, (nc, 3))
    target = rng.random((nc * crop ** 3, 3)).astype(np.float32)
    noise = I.make_noise(nc * crop ** 3, cin, 8, 66)
    loss, out, grads, dg0, dg1 = O.train_forward_backward(grids, params, coord, target, fl, mip, method, noise, size=crop)
    fp = [T(a) for a in grids]
    pt = [T(p) for p in params]
    m = L.make_mlp(pt)
    g = [torch.zeros_like(p) for p in pt]
    gm = L.make_mlp_grad(g)
    g0t, g1t = fp[0], fp[1]
    d0, d1 = torch.zeros_like(g0t), torch.zeros_like(g1t)
    ls = torch.zeros(4, device=dev())
    o = torch.empty((nc * crop ** 3, 3), device=dev())
    geom = L.make_geom(L.METHOD_3D if method == 3 else L.METHOD_3D_V2, g0t, g1t, crop, nc, mip - 2, mip, 6,
                       L.PE_TRIANGULAR if method == 3 else L.PE_SINUSOIDAL)
    h = L.handle(dev())
    coord_t, target_t, noise_t = T(coord, torch.int64), T(target), T(noise)
    for rep in range(2):
        for t in g + [d0, d1, ls]:
            t.zero_()
        L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(g0t), L.ptr(g1t), L.ptr(coord_t), C.byref(m),
                                                   L.ptr(target_t), L.ptr(noise_t), 0, 0, 0, 0, C.byref(gm), L.ptr(d0),
                                                   L.ptr(d1), L.ptr(ls), L.ptr(o), L.PRECISIONS[prec], L.stream_ptr(dev())))
        tol = 1.0 if prec == "f16" else 4.0
        assert abs(float(ls[0]) / (nc * crop ** 3 * 3) - loss) <= 1e-2 * tol * loss
        assert np.abs(o.cpu().numpy() - out).max() <= 4e-3 * tol
        for t, k in zip(g, ("W1", "b1", "W2", "b2", "W3", "b3")):
            assert _rel_l2(t.cpu().numpy(), grads[k]) <= 2e-2 * tol, (k, _rel_l2(t.cpu().numpy(), grads[k]))
        assert _rel_l2(d0.cpu().numpy(), dg0) <= 2e-2 * tol, _rel_l2(d0.cpu().numpy(), dg0)
        assert _rel_l2(d1.cpu().numpy(), dg1) <= 2e-2 * tol, _rel_l2(d1.cpu().numpy(), dg1)


@pytest.mark.parametrize("bits", [8, 4])
def test_decode_from_codes_equals_decode_of_loaded_grids(bits):
    """nic_decode_codes (the quantiser fused into the grid read) is bit-identical to fp_load + decode, on the fast 2-D
    path, the general path (unaligned block) and a 3-D volume."""
    n = nic()
    ic, fpd = n.image_compression, n.fp_def
    size = 256
    configure(IMAGE_SIZE=size)
    grids = I.make_grids(size, 2, bits=bits, seed=98, no_mip=True)
    fp, dec = [T(a) for a in grids], make_decoder(I.make_mlp(73, seed=99, gain=2.0))
    codes = fpd.fp_savable(fp, bits)
    loaded = fpd.fp_load(codes, bits)
    for prec in ("f16", "bf16"):
        a = ic.decode_codes(codes, dec, bits, 0, precision=prec)
        b = ic.decode(loaded, dec, 0, precision=prec, out_dtype=torch.uint8)
        assert torch.equal(a, b)
        a = ic.decode_codes(codes, dec, bits, 0, size=(37, 50), origin=(5, 9), precision=prec)
        b = ic.decode(loaded, dec, 0, size=(37, 50), origin=(5, 9), precision=prec, out_dtype=torch.uint8)
        assert torch.equal(a, b)
    configure(IMAGE_SIZE=32, IMAGE_DIMENSION=3, COMPRESSION_METHOD=3, CROP_MIP_LEVEL=3)
    g3 = [T(a) for a in I.make_grids(32, 3, bits=bits, seed=100, no_mip=True)]
    d3 = make_decoder(I.make_mlp(127, seed=101, gain=2.0))
    c3 = fpd.fp_savable(g3, bits)
    assert torch.equal(ic.decode_codes(c3, d3, bits, 0), ic.decode(fpd.fp_load(c3, bits), d3, 0, precision="f16", out_dtype=torch.uint8))


def test_decode_session_reuses_tables_and_refreshes():
    """DecodeSession: repeated decodes reuse the prepared tables and equal plain decode; after the grids change,
    refresh() (or a different mip level) rebuilds them."""
    n = nic()
    ic = n.image_compression
    size = 256
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    fp = [T(a) for a in I.make_grids(size, 2, seed=102, quantized=True)]
    dec = make_decoder(I.make_mlp(73, seed=103, gain=2.0))
    ses = ic.DecodeSession(fp, dec, precision="f16")
    l0 = n.launch_count(dev())
    a = ses.decode(0)
    l1 = n.launch_count(dev())
    b = ses.decode(0)
    c = ses.decode(0, size=(64, 128), origin=(64, 32))
    l2 = n.launch_count(dev())
    assert torch.equal(a, b) and torch.equal(c, a[64:128, 32:160])
    assert torch.equal(a, ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8))
    assert (l2 - l1) == 2 and (l1 - l0) > 1                    # one kernel per reused call, several for the first
    m1 = ses.decode(1)                                          # another mip level re-keys and rebuilds
    assert torch.equal(m1, ic.decode(fp, dec, 1, precision="f16", out_dtype=torch.uint8))
    with torch.no_grad():
        fp[0].mul_(0.5)
    ses.refresh()
    assert torch.equal(ses.decode(0), ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8))


def test_general_kernel_shapes_against_f32_path():
    """The warp-specialised general decode kernel on unaligned 2-D blocks at several mips, both 3-D methods and
    random-access queries, against the reference-exact fp32 path on the same inputs (north-star tolerance)."""
    n = nic()
    ic = n.image_compression

    def check(fn, prec, n_min=64):
        ref = fn("f32", torch.float32).cpu().numpy()
        out = fn(prec, torch.float32).cpu().numpy()
        u8 = fn(prec, torch.uint8).cpu().numpy()
        assert np.abs(out - ref).max() < (4e-3 if prec == "f16" else 2.5e-2)
        within1, same, worst = lsb_stats(u8, O.quantize_to_bit(ref, 8).astype(np.uint8))
        assert worst <= (1 if prec == "f16" else 2), (prec, worst)
        if ref.size >= n_min:
            assert within1 >= 0.999, (prec, within1, same)
        assert np.array_equal(u8, np.floor(out * np.float32(255) + np.float32(0.5)).astype(np.uint8))

    size = 256
    configure(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    fp = [T(a) for a in I.make_grids(size, 2, seed=104, quantized=True)]
    dec = make_decoder(I.make_mlp(73, seed=105, gain=2.0))
    for prec in ("f16", "bf16"):
        for mip, sz, org in ((0, (37, 50), (5, 9)), (0, (250, 3), (1, 200)), (1, (128, 128), (0, 0)), (3, (32, 32), (0, 0)), (5, (8, 8), (0, 0))):
            check(lambda p_, dt: ic.decode(fp, dec, mip, size=sz, origin=org, precision=p_, out_dtype=dt), prec)
    for method, cin in ((3, 127), (4, 79)):
        configure(IMAGE_SIZE=32, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=3)
        g3 = [T(a) for a in I.make_grids(32, 3, seed=106, no_mip=True, quantized=True)]
        d3 = make_decoder(I.make_mlp(cin, seed=107, gain=2.0))
        check(lambda p_, dt: ic.decode(g3, d3, 0, precision=p_, out_dtype=dt), "f16")
        check(lambda p_, dt: ic.decode(g3, d3, 0, size=(5, 7, 3), origin=(9, 2, 20), precision=p_, out_dtype=dt), "bf16")
        q = torch.tensor(np.random.default_rng(108).integers(0, 32, (1003, 3)))
        check(lambda p_, dt: ic.decode_points(g3, d3, q, 0, precision=p_, out_dtype=dt), "f16")


def test_end_to_end_training_psnr_vs_reference_port():
    """End to end, like the reference's own validation (final PSNR, image_compression.py:482-489): 120 training steps on
    a 256^2 synthetic image with one full-frame crop per step, (a) by the torch-CPU port of the reference loop (autograd +
    torch.optim.Adam + CosineAnnealingLR + clamp, its own torch noise stream), (b) by FusedTrainer fp32 and (c) f16 with
    in-kernel Philox noise.  Same initial grids and decoder; the noise streams differ, so the three runs are compared
    statistically: final full-frame PSNR (reference formula) within 0.3 dB of the port, and the two GPU paths within
    0.1 dB of each other."""
    from oracle import nic_oracle_torch as OT
    n = nic()
    ic = n.image_compression
    size, steps = 256, 120
    img = I.make_image(size, 2, seed=110)
    grids0 = I.make_grids(size, 2, seed=111, no_mip=True)
    params0 = I.make_mlp(73, seed=112)
    table = O.create_pyramid_mip_levels(size, size // 4)
    coord = np.array([[0, 0]])
    target = img.reshape(3, -1).T[None]
    target8 = np.floor(img.transpose(1, 2, 0) * 255 + 0.5)

    def psnr_of(decoded01):
        return O.calculate_psnr(target8.astype(np.float32), O.quantize_to_bit(decoded01, 8).astype(np.float32))

    # (a) reference port on the host
    torch.manual_seed(7)
    ref_dec = OT.make_decoder(params0)
    tr = OT.Trainer(grids0, ref_dec, steps, 8, 1, table)
    for _ in range(steps):
        tr.step(coord, target, 0)
    gq = [torch.tensor(O.quantize4fp(g.detach().numpy(), 8)) for g in tr.fp]
    psnr_ref = psnr_of(OT.decode_block(gq, ref_dec, size, 0, table, 1).numpy())
    # (b), (c) fused trainers
    psnr = {}
    for prec in ("f32", "f16"):
        configure(IMAGE_SIZE=size)
        fp = [T(a) for a in grids0]
        dec = make_decoder(params0)
        ft = ic.FusedTrainer(fp, dec, num_epochs=steps, fp_bits=8, seed=3, precision=prec)
        tg = T(target)
        for _ in range(steps):
            ft.step(torch.tensor(coord), tg, 0)
        out = ic.decode(n.fp_def.fp_all_quantize(ft.fp, 8), dec, 0, precision="f32").cpu().numpy()
        psnr[prec] = psnr_of(out)
    assert psnr_ref > 18.0, psnr_ref
    assert abs(psnr["f32"] - psnr_ref) <= 0.3 and abs(psnr["f16"] - psnr_ref) <= 0.3, (psnr_ref, psnr)
    assert abs(psnr["f16"] - psnr["f32"]) <= 0.1, psnr


def test_edge_cases_empty_and_oversized():
    """Empty inputs are no-ops on every fused entry point; sample counts beyond the kernels' 32-bit index are refused
    with NIC_ERR_UNSUPPORTED (never a silent wrap), checked before anything is launched or dereferenced."""
    import ctypes as C
    n = nic()
    ic, L = n.image_compression, n._lib
    size = 64
    configure(IMAGE_SIZE=size)
    fp = [T(a) for a in I.make_grids(size, 2, seed=113, no_mip=True, quantized=True)]
    pt = [T(p) for p in I.make_mlp(73, seed=114)]
    dec = make_decoder([p.cpu().numpy() for p in pt])
    h, lib = L.handle(dev()), L.load_library()
    # empty decode / empty query batch / empty crop batch
    for prec in ("f32", "f16"):
        assert ic.decode(fp, dec, 0, size=(0, 16), precision=prec).shape == (0, 16, 3)
        assert ic.decode_points(fp, dec, torch.zeros((0, 2), dtype=torch.int64), 0, precision=prec).shape == (0, 3)
    m = L.make_mlp(pt)
    g = [torch.zeros_like(p) for p in pt]
    gm = L.make_mlp_grad(g)
    d0, d1, ls = torch.zeros_like(fp[0]), torch.zeros_like(fp[1]), torch.zeros(4, device=dev())
    dummy = torch.zeros((1, 2), dtype=torch.int64, device=dev())
    for prec in (L.PREC_F32, L.PREC_F16):
        geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], 16, 0, -2, 0, 6, L.PE_TRIANGULAR)
        L.check(h, lib.nic_train_step(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), L.ptr(dummy), C.byref(m), None, None, 0, 0, 0, 0,
                                      C.byref(gm), L.ptr(d0), L.ptr(d1), L.ptr(ls), None, prec, L.stream_ptr(dev())))
    torch.cuda.synchronize()
    assert float(ls[0]) == 0 and all(float(t.abs().max()) == 0 for t in g + [d0, d1])
    # 2^32 samples: refused
    geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], (1024, 1024), 4096, -2, 0, 6, L.PE_TRIANGULAR)
    out = torch.zeros(16, dtype=torch.uint8, device=dev())
    for prec in (L.PREC_F16, L.PREC_BF16):
        rc = lib.nic_decode(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), L.ptr(dummy), C.byref(m), L.ptr(out), L.DT_U8, prec,
                            L.stream_ptr(dev()))
        assert rc == -2, rc
    tg = torch.zeros(16, device=dev())
    rc = lib.nic_train_step(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), L.ptr(dummy), C.byref(m), L.ptr(tg), None, 0, 0, 0, 0,
                            C.byref(gm), L.ptr(d0), L.ptr(d1), L.ptr(ls), None, L.PREC_F16, L.stream_ptr(dev()))
    assert rc == -2, rc
    torch.cuda.synchronize()


def test_exchange_kernel_single_rank_equals_plain_adam():
    """nic_adam_step_exchange with world = 1 (runs on a 1-GPU box): symmetric buffer allocation and aliasing, the flag
    protocol against itself (one-shot and sliced: both handshakes, the grid-wide arrive counter, the write-back), the
    other-parity clear and the loss bookkeeping — and bit-equality with nic_adam_step_loss on the same gradients (a
    one-term sum is exact).  The multi-rank behaviour is test_dp_exchange_* below."""
    n = nic()
    L = n._lib
    import ctypes as C
    rng = np.random.default_rng(5)
    sizes = [12 * 33 * 33, 73 * 64, 64, 4]          # a grid, W1, b1, the loss slot
    offs, total = [], 0
    for sz in sizes:
        offs.append(total)
        total += (sz + 3) // 4 * 4
    buf = L.SymmetricBuffer(dev(), 2 * total + 64)
    try:
        assert buf.tensor.data_ptr() == buf.ptr and float(buf.tensor.abs().sum()) == 0.0
        flat0, flat1 = buf.tensor[:total], buf.tensor[total:2 * total]
        grad = T(rng.standard_normal(total).astype(np.float32) * 1e-2)
        results = {}
        for mode in ("plain", "exchange", "sliced"):
            params = [T(rng2) for rng2 in (np.random.default_rng(6).standard_normal(sz).astype(np.float32) * 0.1 for sz in sizes[:3])]
            ms = [torch.full_like(p, 0.01) for p in params]
            vs = [torch.full_like(p, 0.001) for p in params]
            g = grad.clone() if mode == "plain" else flat0
            if mode != "plain":
                flat0.copy_(grad)
                flat1.fill_(7.0)                     # stale contents of the other parity: must come back cleared
            arr = (L.NicAdamTensor * 3)()
            for k in range(3):
                a = arr[k]
                a.p, a.g = params[k].data_ptr(), g[offs[k]:].data_ptr()
                a.m, a.v, a.numel = ms[k].data_ptr(), vs[k].data_ptr(), sizes[k]
                a.lr, a.t, a.clamp, a.clamp_lo, a.clamp_hi = 0.01, 3, int(k == 0), -0.05, 0.05
            loss = torch.zeros(1, device=dev())
            loss_sum = g[offs[3]:offs[3] + 4]
            h, lib = L.handle(dev()), L.load_library()
            if mode == "plain":
                L.check(h, lib.nic_adam_step_loss(h, arr, 3, 0.9, 0.999, 1e-8, 1.0, 0, L.ptr(loss_sum), L.ptr(loss), 0.5,
                                                  L.stream_ptr(dev())))
            else:
                x = L.NicExchange()
                x.world, x.rank, x.token = 1, 0, 41 if mode == "exchange" else 42
                x.reserved = L.EXCHANGE_SLICED if mode == "sliced" else L.EXCHANGE_ONE_SHOT
                x.peer_flat[0], x.peer_flag[0] = buf.ptr, buf.ptr + 8 * total
                x.zero_buf, x.zero_numel = buf.ptr + 4 * total, total
                L.check(h, lib.nic_adam_step_exchange(h, arr, 3, 0.9, 0.999, 1e-8, 1.0, C.byref(x), L.ptr(loss_sum), L.ptr(loss),
                                                      0.5, L.stream_ptr(dev())))
                assert not L.exchange_status(dev())
                assert float(flat1.abs().sum()) == 0.0
                flags = buf.tensor[2 * total:2 * total + 32].view(torch.int32)
                assert int(flags[0]) == x.token                                                   # slot 0 of the flag array
                if mode == "sliced":
                    assert int(flags[16]) == 42                                                   # ... and of the second handshake
                assert torch.equal(flat0, grad)       # one-shot leaves the buffer alone; a one-rank slice sum rewrites the same bits
            results[mode] = [t.cpu().numpy() for t in params + ms + vs] + [loss.cpu().numpy()]
        for a, b, c in zip(results["plain"], results["exchange"], results["sliced"]):
            assert np.array_equal(a, b) and np.array_equal(a, c)
        assert abs(float(results["plain"][-1][0]) - 0.5 * float(grad[offs[3]])) < 1e-7
    finally:
        buf.close()


def test_dp_exchange_peer_memory_matches_nccl():
    """Data parallel training on 2 GPUs (skipped on a 1-GPU box): the exchange step fused into the optimiser over NVLink
    peer memory (nic_adam_step_exchange) against the NCCL all-reduce — bit-identical replicas across ranks in both modes,
    the two modes equal to rounding, no timeouts.  The check itself lives in tests/dp_exchange_check.py (torchrun)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "dp_exchange_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

