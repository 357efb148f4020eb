"""The BASELINE.json configurations through the public Task API on the B200,
against the oracle (full size where the oracle finishes in seconds, otherwise
reduced in duration only) and through size-independent properties.
"""
import numpy as np
import pytest

import bbt_oracle as orc

from test_kernels import assert_voltage, assert_power, cnoise, rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def bt():
    import baseband_tasks_b200 as pkg
    return pkg


def t0(bt):
    return bt.Time(1289567655)


def test_c1_dedisperse_full_size(bt):
    """configs[0]: 1-ch complex64 16 MHz, 2^22 samples, DM=26.8, N=2^21."""
    n, rate, freq, dm, spf = 1 << 22, 16e6, 400e6, 26.8, 1206812
    src = bt.NoiseGenerator((n,), t0(bt), rate, samples_per_frame=1 << 20,
                            dtype='c8', seed=1234568, frequency=freq,
                            sideband=1)
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    assert dd._ih_samples_per_frame == 1 << 21
    assert (dd._pad_start, dd._pad_end) == (431817, 458523)
    assert dd.shape == (3303964,)
    x = orc.noise_stream(1234568, n, 1 << 20)
    op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, n, 1 << 20,
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    want = orc.disperse(x, op)            # 3 frames, the last re-anchored
    got = dd.read()
    assert_voltage(got, want)
    # In double precision the oracle differs from itself by about as much.
    dd.seek(2 * spf - 5)
    assert_voltage(dd.read(10), want[2 * spf - 5:2 * spf + 5])


@pytest.mark.parametrize('step', [1e-3, 8])
def test_c2_chain(bt, step):
    """configs[1], 3 frames: 8ch x 2pol -> Dedisperse(100) -> Channelize(1024)
    -> Power -> Integrate(1 ms, or 8 spectra for the tie-free check)."""
    rate, dm, log2n = 8e6, 100., 20
    freq = (1372e6 + 8e6 * np.arange(8)).reshape(8, 1)
    N = 1 << log2n
    pad = 74847 + 80161
    spf = N - pad
    n = 2 * spf + N
    shape = (8, 2)
    src = bt.NoiseGenerator((n,) + shape, t0(bt), rate,
                            samples_per_frame=1 << 18, dtype='c8',
                            seed=1234569, frequency=freq, sideband=1,
                            polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    assert (dd._pad_start, dd._pad_end) == (74847, 80161)
    assert dd._ih_samples_per_frame == N and dd.shape[0] == 3 * spf
    it = bt.Integrate(bt.Power(bt.Channelize(dd, 1024)), step, average=False)
    assert it._fused == 'chanpow'
    got = it.read()
    x = orc.noise_stream(1234569, n, 1 << 18, shape)
    op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, n, 1 << 18,
                          shape, samples_per_frame=spf,
                          fast_len=orc.next_pow2)
    y = orc.disperse(x, op)
    dd.seek(spf - 100)
    assert_voltage(dd.read(200), y[spf - 100:spf + 100])
    power = orc.power(orc.channelize(y, 1024), axis=-1)
    ip = orc.IntegratePlan(power.shape[0], rate / 1024, step)
    offsets = ip.offsets(np.arange(ip.n_out + 1))
    want, count = orc.integrate(power, offsets)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got['count'][:, 0, 0, 0], count.ravel())
    assert_power(got['data'], want)
    avg = bt.Integrate(it.ih, step).read()
    assert_power(avg, want / count)
    if step == 8:
        assert np.all(count == 8)
    else:
        assert set(np.unique(count)) == {7, 8}


def test_c3_pfb_dedisperse_power(bt):
    """configs[2], reduced in duration: real 8-bit-valued 800 MS/s, 2 pol ->
    4-tap 1024(+1)-channel PFB -> per-channel Dedisperse -> Power."""
    n_chan_in, n_tap = 2048, 4
    n_spec_in = 2200
    n = n_spec_in * n_chan_in
    rate = 800e6
    rng = np.random.default_rng(1234570)
    x = np.clip(np.round(rng.normal(size=(n, 2)) * 20), -127, 127).astype('f4')
    src = bt.ArrayStream(x, t0(bt), rate, samples_per_frame=1 << 16,
                         frequency=800e6, sideband=-1,
                         polarization=np.array(['X', 'Y']))
    response = bt.sinc_hamming(n_tap, n_chan_in)
    pfb = bt.PolyphaseFilterBank(src, response)
    assert pfb.shape[1:] == (1025, 2) and pfb.sample_rate == rate / 2048
    dm = 10.
    dd = bt.Dedisperse(pfb, dm, reference_frequency=pfb.frequency)
    pw = bt.Power(dd)
    want_pfb = orc.pfb(x.astype('f8'), response, ih_samples_per_frame=1 << 16)
    assert pfb.shape == want_pfb.shape
    assert_voltage(pfb.read(), want_pfb.astype('c8'))
    freq = orc.channelize_frequency(800., -1, 2048, rate / 1e6, True, 1)
    op = orc.DispersePlan(-dm, freq, -1, rate / 2048 / 1e6, True,
                          want_pfb.shape[0], pfb.samples_per_frame, (1025, 2),
                          reference_frequency_mhz=freq,
                          fast_len=orc.next_pow2)
    assert (dd._pad_start, dd._pad_end, dd._ih_samples_per_frame) == (
        op.pad_start, op.pad_end, op.N)
    y = orc.disperse(want_pfb.astype('c8'), op)
    assert_voltage(dd.read(), y)
    assert_power(pw.read(), orc.power(y, axis=-1))


def test_c4_frame_full_size(bt):
    """configs[3]: 512 MHz dual-pol, DM=1000, 2^24-point frames: one and a
    half frames against the oracle, plus linearity at full size."""
    rate, freq, dm = 512e6, 8192e6, 1000.
    N = 1 << 24
    pad_start, pad_end = 1889551, 2075345
    spf = N - pad_start - pad_end
    n = N + spf // 2
    rng = np.random.default_rng(1234571)
    x = cnoise(rng, (n, 2))
    src = bt.ArrayStream(x, t0(bt), rate, samples_per_frame=1 << 20,
                         frequency=freq, sideband=1,
                         polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    assert (dd._pad_start, dd._pad_end) == (pad_start, pad_end)
    assert dd._ih_samples_per_frame == N
    got = dd.read()
    op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, n, 1 << 20,
                          (2,), samples_per_frame=spf,
                          fast_len=orc.next_pow2)
    want = orc.disperse(x, op)
    assert got.shape == want.shape == (n - pad_start - pad_end, 2)
    assert_voltage(got, want)
    # Linearity: D(a x + b x') = a D(x) + b D(x').
    x2 = cnoise(rng, (n, 2))
    a, b = 0.75 - 0.5j, -1.25 + 0.3j
    mix = bt.ArrayStream((a * x + b * x2).astype('c8'), t0(bt), rate,
                         frequency=freq, sideband=1)
    other = bt.ArrayStream(x2, t0(bt), rate, frequency=freq, sideband=1)
    d_mix = bt.Dedisperse(mix, dm, samples_per_frame=spf).read(spf)
    d_other = bt.Dedisperse(other, dm, samples_per_frame=spf).read(spf)
    # (Three results, each within 1e-5 of the oracle, are combined here:
    # the bound for the combination is (1 + |a| + |b|) 1e-5 = 3.2e-5.)
    assert_voltage(d_mix, a * got[:spf] + b * d_other, tol=2e-5)
    # Energy: the chirp is a pure phase, so a whole frame keeps its power
    # (valid part against the matching part of the circular result).
    assert abs(rms(got[:spf]) / rms(x[pad_start:pad_start + spf]) - 1) < 1e-3


def test_c5_fold_full_frame(bt):
    """configs[4]: Dedisperse -> Power -> Fold(512, polynomial) over one
    2^24-point frame against the oracle chain (disperse -> power -> fold):
    counts bit-exact, sums to 1e-5."""
    rate, freq, dm = 512e6, 8192e6, 1000.
    N = 1 << 24
    pad_start = 1889551
    spf = N - pad_start - 2075345
    rng = np.random.default_rng(1234572)
    x = cnoise(rng, (N, 2))
    src = bt.ArrayStream(x, t0(bt), rate, frequency=freq, sideband=1,
                         polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, N, N, (2,),
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    power = orc.power(orc.disperse(x, op), axis=-1)
    assert power.shape == (spf, 4)
    poly = bt.PolynomialPhase([0.25, 29.946923, -3.77535e-10 / 2.], t0(bt))
    # A pulsar this slow fills 2 of 512 bins in 25 ms: spin it up so that
    # all bins are visited and bin edges are crossed ~10^5 times.
    fast = bt.PolynomialPhase([0.25, 29.946923e3, -3.77535e-4 / 2.], t0(bt))
    for phase in (poly, fast):
        fold = bt.Fold(bt.Power(dd), 512, phase, average=False)
        assert fold._fused == 'power'
        # Samples are counted on the grid of the source stream: the first
        # dedispersed sample is its sample pad_start.
        i_ref, i_0 = phase.grid(fold.ih)
        assert (i_ref, i_0) == (0., pad_start)
        got = fold.read()
        want, wcount = orc.fold(      # sums in float64
            power.astype('f8'), np.array([0, spf]), 512,
            lambda i: phase.of_index(i + i_0, i_ref, rate))
        np.testing.assert_array_equal(got['count'],
                                      np.broadcast_to(wcount, got.shape))
        assert got['count'][..., 0].sum() == spf
        assert_power(got['data'], want.astype('f4'))


def test_empty_and_edge_reads(bt):
    """Zero-length reads, reads to exactly the end, and one past it."""
    x = cnoise(np.random.default_rng(3), (3 * 4096, 2))
    src = bt.ArrayStream(x, t0(bt), 1e6, frequency=300e6, sideband=1,
                         polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, 3., samples_per_frame=4096 - 923)
    assert dd.read(0).shape == (0, 2)
    dd.seek(0, 2)
    assert dd.read().shape == (0, 2)
    with pytest.raises(EOFError):
        dd.read(1)
    pw = bt.Power(bt.Channelize(dd, 64))
    pw.seek(-1, 2)
    assert pw.read().shape == (1, 64, 4)
    with pytest.raises(AssertionError):
        bt.Channelize(dd, 1 << 15)     # frame larger than the stream


@pytest.mark.parametrize('which', ['C4', 'C2'])
def test_bench_block_shape(bt, which):
    """The block shape bench.py launches (many frames in ONE launch of every
    kernel): 8 frames of 2^24 x 2 series (C4), 16 frames of 2^20 x 16 series
    (C2).  The first, a middle and the last frame of the dedispersed block
    against the oracle, and the integrated spectra of those frames."""
    if which == 'C4':
        rate, freq, dm, N, n_frames = 512e6, 8192e6, 1000., 1 << 24, 8
        shape, pads = (2,), (1889551, 2075345)
    else:
        rate, dm, N, n_frames = 8e6, 100., 1 << 20, 16
        freq = (1372e6 + 8e6 * np.arange(8)).reshape(8, 1)
        shape, pads = (8, 2), (74847, 80161)
    spf = N - sum(pads)
    n = (n_frames - 1) * spf + N
    g = np.random.default_rng(4242)
    # float32 draws directly: the block is 1.7 GB (C4).
    x = np.empty((n,) + shape, np.complex64)
    xv = x.view(np.float32)
    step = 1 << 22
    for i in range(0, n, step):
        xv[i:i + step] = g.standard_normal(xv[i:i + step].shape,
                                           dtype=np.float32)
    src = bt.ArrayStream(bt._buffers.as_device(x), t0(bt), rate,
                         samples_per_frame=1 << 20, frequency=freq,
                         sideband=1, polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    assert (dd._pad_start, dd._pad_end) == pads
    assert dd._ih_samples_per_frame == N and dd.shape[0] == n_frames * spf
    dd.read_device(1)                    # creates the plan (chirp kernel)
    dd.seek(0)
    launches0 = bt._cabi.lib().bbt_launch_count()
    y = dd.read_device()                 # all frames, one launch per pass
    assert bt._cabi.lib().bbt_launch_count() - launches0 == 3
    it = bt.Integrate(bt.Power(bt.Channelize(dd, 1024)), 8, average=False)
    spectra = it.read()
    for f in (0, n_frames // 2, n_frames - 1):
        xf = x[f * spf:f * spf + N]
        op = orc.DispersePlan(-dm, np.asarray(freq) / 1e6, 1, rate / 1e6, True,
                              N, N, shape, samples_per_frame=spf,
                              fast_len=orc.next_pow2)
        want = orc.disperse(xf, op)
        assert want.shape[0] == spf
        got = y[f * spf:(f + 1) * spf].cpu().numpy()
        assert_voltage(got, want)
        # Integration bins of 8 spectra that lie wholly inside this frame.
        b0 = -(-f * spf // (8 * 1024))
        b1 = (f + 1) * spf // (8 * 1024)
        lo = b0 * 8 * 1024 - f * spf
        power = orc.power(orc.channelize(want[lo:lo + (b1 - b0) * 8192],
                                         1024), axis=-1)
        wsum = power.reshape((b1 - b0, 8) + power.shape[1:]).sum(1)
        assert np.all(spectra['count'][b0:b1] == 8)
        assert_power(spectra['data'][b0:b1], wsum)


def test_c4_time_block_sharding(bt):
    """configs[3] shared out in time: two ranks' blocks (frames plus the
    overlap-save halo, `parallel.StreamBlock`) processed one after the other
    on this GPU give, after adding the partial sums of the bin the cut runs
    through, exactly the counts and (to 1e-5) the sums of the whole stream;
    with Fold the phase bins are bit-identical."""
    from baseband_tasks_b200 import parallel
    rate, freq, dm, N = 512e6, 8192e6, 1000., 1 << 24
    pads = (1889551, 2075345)
    spf, pad = N - sum(pads), sum(pads)
    n_frames, world = 3, 2
    n = n_frames * spf + pad
    g = np.random.default_rng(777)
    x = np.empty((n, 2), np.complex64)
    xv = x.view(np.float32)
    for i in range(0, n, 1 << 22):
        xv[i:i + (1 << 22)] = g.standard_normal(xv[i:i + (1 << 22)].shape,
                                                dtype=np.float32)
    kw = dict(frequency=freq, sideband=1, polarization=np.array(['X', 'Y']))

    def chain(src, fold):
        dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
        if fold:
            poly = bt.PolynomialPhase([0.25, 29.946923e3, -3.77535e-4 / 2.],
                                      t0(bt))
            return bt.Fold(bt.Power(dd), 512, poly, average=False)
        return bt.Integrate(bt.Power(bt.Channelize(dd, 1024)), 1e-3,
                            average=False)

    whole = bt.ArrayStream(bt._buffers.as_device(x), t0(bt), rate,
                           samples_per_frame=1 << 20, **kw)
    for fold in (False, True):
        unit = 1 if fold else 1024
        ref = chain(whole, fold)
        plans = [parallel.block_plan(n_frames, spf, pad, unit, r, world)
                 for r in range(world)]
        assert plans[0][1] == plans[1][0] and plans[0][0] == 0
        total_s = total_c = None
        for first, last, in0, in1 in plans:
            block = parallel.StreamBlock(
                bt._buffers.as_device(x[in0:in1]), in0, n, t0(bt), rate,
                samples_per_frame=1 << 20, **kw)
            it = chain(block, fold)
            assert it.shape == ref.shape
            if fold:
                it.seek(0)
                s, c = it.read_sums(within=(first, last))
            else:
                edges = it._get_offsets(np.arange(it.shape[0] + 1))
                b0, b1 = parallel.bin_range(edges, first // unit, last // unit)
                it.seek(b0)
                part = it.read_sums(b1 - b0, within=(first // unit,
                                                     last // unit))
                s = part[0].new_zeros((it.shape[0],) + part[0].shape[1:])
                c = part[1].new_zeros((it.shape[0],) + part[1].shape[1:])
                s[b0:b1], c[b0:b1] = part
            total_s = s if total_s is None else total_s + s
            total_c = c if total_c is None else total_c + c
        ref.seek(0)
        n_whole = plans[-1][1] // unit          # complete frames only
        if fold:
            rs, rc = ref.read_sums(within=(0, plans[-1][1]))
        else:
            rs, rc = ref.read_sums()
            edges = ref._get_offsets(np.arange(ref.shape[0] + 1))
            full = edges[1:] <= n_whole
            rs, rc = rs[full], rc[full]
            total_s, total_c = total_s[full], total_c[full]
        np.testing.assert_array_equal(total_c.cpu().numpy(),
                                      rc.cpu().numpy())
        assert_power(total_s.cpu().numpy(), rs.cpu().numpy())


def test_c4_power_fused_planar(bt):
    """Power straight after the 2^24-point dedispersion of configs[3] (planar
    work buffer: the polarizations of a pair are eight lanes apart in the last
    pass) equals Power of the dedispersed voltages."""
    rate, freq, dm, N = 512e6, 8192e6, 1000., 1 << 24
    spf = N - 1889551 - 2075345
    n = spf + N + 12345
    g = np.random.default_rng(778)
    x = np.empty((n, 2), np.complex64)
    xv = x.view(np.float32)
    for i in range(0, n, 1 << 22):
        xv[i:i + (1 << 22)] = g.standard_normal(xv[i:i + (1 << 22)].shape,
                                                dtype=np.float32)
    kw = dict(frequency=freq, sideband=1, polarization=np.array(['X', 'Y']))
    dev = bt._buffers.as_device(x)

    def chain():
        src = bt.ArrayStream(dev, t0(bt), rate, samples_per_frame=1 << 20,
                             **kw)
        return bt.Dedisperse(src, dm, samples_per_frame=spf)

    fused = bt.Power(chain())
    assert fused._fused
    dd = chain()
    for start, count in ((0, 2 * spf), (spf - 1000, 5000),
                         (2 * spf, fused.shape[0] - 2 * spf)):
        dd.seek(start)
        v = dd.read(count)
        want = orc.power(v.astype(np.complex128), axis=1)
        fused.seek(start)
        got = fused.read(count)
        assert got.shape == want.shape and got.dtype == np.float32
        assert_power(got, want)
