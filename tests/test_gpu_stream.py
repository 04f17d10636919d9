"""GPU parity: streaming front-end and CS16 conversion through the C ABI.
Block k's payload must be what the k-th DeModulateBytes call returns on the oracle (TB/SDR/ModDemodOverSDR.cs:127-136)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

START, END = b"MESSAGE_START", b"MESSAGE_STOP"
TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"   # testAtDataLevel.cs:20-22
FS = 10_000_000
RS = FS // 2
ALPHA = float(np.float32(0.4))
SPAN = 10


def _bursts(orc, n=8, seed=3):
    """Framed random payloads, one burst per ModulateBytes call, through the two-unstable-LO channel
    (testAtDataLevel.cs:24-42).  With the TSC in place every burst after the first decodes when each burst is one
    DeModulateBytes call — the per-call semantics the stream has to keep."""
    mod = orc.QPSKModulator(FS, RS, ALPHA, SPAN, True, TSC)
    rng = np.random.default_rng(seed)
    pays = [rng.integers(0, 256, int(rng.integers(5, 400)), dtype=np.uint8).tobytes() for _ in range(n)]
    tx, rx = orc.NCO(100e6, FS, 1, seed=seed, stream=0), orc.NCO(100e6, FS, 1, seed=seed, stream=1)
    return pays, [orc.channel_apply(tx, rx, 0, mod.ModulateBytes(p, START, END)) for p in pays]


def _oracle(orc):
    return orc.QPSKDeModulator(FS, RS, ALPHA, SPAN, tsc=TSC)


def _gpu_demod(gpu):
    d = gpu.QPSKDeModulator(FS, RS, ALPHA, SPAN, tsc=TSC)
    d.set_fir_mode(gpu.FIR_EXACT)
    return d


@pytest.mark.parametrize("depth", [1, 3, 8])
def test_stream_one_burst_per_block_equals_per_call_oracle(gpu, orc, depth):
    pays, bursts = _bursts(orc)
    od = _oracle(orc)
    want = [od.DeModulateBytes(b, START, END, cap=1 << 16) for b in bursts]
    assert want[1:] == pays[1:]                                    # every burst after the first decodes
    st = gpu.StreamingDemodulator(_gpu_demod(gpu), START, END, max_block_floats=max(b.size for b in bursts),
                                  max_payload_bytes=1 << 16, depth=depth)
    got = []
    for i, b in enumerate(bursts):
        st.push(b)
        if i % 3 == 2:                                             # poll now and then, like a consumer thread would
            while True:
                p = st.poll()
                if p is None:
                    break
                got.append(p)
    got += st.drain()
    assert got == want
    assert st.pending() == 0 and st.poll() is None
    st.close()


def test_stream_with_fll_time_chunk_pipeline(gpu, orc):
    """FLL opt-in under the front-end: every block is long enough (>= 2048 samples) for the demodulator's time-chunk
    pipeline, whose side stream and events then run beneath the front-end's copy / compute streams.  Payload of block k
    = the oracle's k-th DeModulateBytes call with the FLL."""
    mod = orc.QPSKModulator(FS, RS, ALPHA, SPAN, True, TSC)
    rng = np.random.default_rng(12)
    pays = [rng.integers(0, 256, int(rng.integers(300, 700)), dtype=np.uint8).tobytes() for _ in range(7)]
    tx, rx = orc.NCO(100e6, FS, 1, seed=12, stream=0), orc.NCO(100e6, FS, 1, seed=12, stream=1)
    bursts = [orc.channel_apply(tx, rx, 0, mod.ModulateBytes(p, START, END)) for p in pays]
    assert min(b.size for b in bursts) >= 2 * 2048
    od = orc.QPSKDeModulator(FS, RS, ALPHA, SPAN, tsc=TSC, use_fll=True)
    want = [od.DeModulateBytes(b, START, END, cap=1 << 16) for b in bursts]
    gd = gpu.QPSKDeModulator(FS, RS, ALPHA, SPAN, tsc=TSC, use_fll=True)
    gd.set_fir_mode(gpu.FIR_EXACT)
    st = gpu.StreamingDemodulator(gd, START, END, max_block_floats=max(b.size for b in bursts), max_payload_bytes=1 << 16, depth=3)
    got = []
    for b in bursts:
        st.push(b)
        p = st.poll()
        if p is not None:
            got.append(p)
    got += st.drain()
    assert got == want
    st.close()


@pytest.mark.parametrize("mtu", [2040, 1000])
def test_stream_mtu_blocks_equal_per_call_oracle(gpu, orc, mtu):
    """A continuous stream cut at radio-MTU boundaries (ModDemodOverSDR.cs:127-136): whatever each per-block call
    returns on the oracle — mostly nothing, the TSC strip is per call — the stream returns for that block."""
    _, bursts = _bursts(orc, seed=4)
    y = np.concatenate(bursts)
    blocks = [y[a:a + 2 * mtu] for a in range(0, y.size, 2 * mtu)]
    od = _oracle(orc)
    want = [od.DeModulateBytes(b, START, END, cap=1 << 16) for b in blocks]
    st = gpu.StreamingDemodulator(_gpu_demod(gpu), START, END, max_block_floats=2 * mtu, max_payload_bytes=1 << 16, depth=4)
    for b in blocks:
        st.push(b)
    assert st.drain() == want


def test_stream_spill_when_nobody_polls(gpu, orc):
    """More blocks pushed than slots without a single poll: results are parked and still come out in order."""
    pays, bursts = _bursts(orc, n=9, seed=8)
    od = _oracle(orc)
    want = [od.DeModulateBytes(b, START, END, cap=1 << 16) for b in bursts]
    st = gpu.StreamingDemodulator(_gpu_demod(gpu), START, END, max_block_floats=max(b.size for b in bursts),
                                  max_payload_bytes=1 << 16, depth=2)
    for b in bursts:
        st.push(b)
    assert st.pending() == len(bursts)
    assert st.drain() == want
    assert sum(1 for w in want if w) >= 2


def test_stream_cs16_ingest(gpu, orc):
    """CS16 blocks: identical to the oracle on the widened cf32 blocks."""
    pays, bursts = _bursts(orc, seed=5)
    scale = float(np.float32(1.0 / 32768.0))
    peak = max(np.abs(b).max() for b in bursts)
    b16 = [np.clip(np.round(b / peak * 30000.0), -32768, 32767).astype(np.int16) for b in bursts]
    bf = [(q.astype(np.float32) * np.float32(scale)).astype(np.float32) for q in b16]
    od = _oracle(orc)
    want = [od.DeModulateBytes(b, START, END, cap=1 << 16) for b in bf]
    assert sum(1 for w in want if w) >= 5
    st = gpu.StreamingDemodulator(_gpu_demod(gpu), START, END, max_block_floats=max(b.size for b in bursts),
                                  max_payload_bytes=1 << 16, depth=3)
    for q in b16:
        st.push_cs16(q, scale)
    assert st.drain() == want
    assert np.array_equal(gpu.Cs16ToCf32(b16[0], scale), bf[0])


def test_stream_errors(gpu, orc):
    gd = gpu.QPSKDeModulator(FS, RS, ALPHA, SPAN)
    with pytest.raises(gpu.ArgumentException):
        gpu.StreamingDemodulator(gd, b"", END, 4096)              # empty marker (QPSKDeModulator.cs:174-175)
    st = gpu.StreamingDemodulator(gd, START, END, max_block_floats=4096, max_payload_bytes=64, depth=2)
    with pytest.raises(gpu.ArgumentException):
        st.push(np.zeros(5, np.float32))                           # odd interleaved length
    with pytest.raises(gpu.QpskCudaError):
        st.push(np.zeros(4098, np.float32))                        # larger than max_block_floats
    st.push(np.zeros(0, np.float32))                               # empty block = empty call (:350-351)
    assert st.drain() == [b""]
    batch = gpu.QPSKDeModulator(FS, RS, ALPHA, SPAN, channels=2)
    with pytest.raises(gpu.QpskCudaError):
        gpu.StreamingDemodulator(batch, START, END, 4096)


def test_save_as_cs16_matches_reference_semantics(gpu, orc):
    from oracle import np_twin
    rng = np.random.default_rng(11)
    for x in (rng.standard_normal(2 * 5000).astype(np.float32) * np.float32(0.37),
              np.array([1.0, -1.0, 0.5, -0.5, 0.99999, -0.99999, 3.0517578e-05, -3.0517578e-05], np.float32),
              np.zeros(16, np.float32),                              # maxVal < 1e-12 -> 1.0 (:92)
              np.array([1e-13, -1e-13], np.float32),
              np.array([-2.5, 0.1], np.float32)):
        want, wmax = np_twin.save_as_cs16(x)
        got, gmax = gpu.SaveAsCs16(x)
        assert np.array_equal(got, want)
        assert np.float32(gmax) == np.float32(wmax)
    with pytest.raises(gpu.ArgumentException):
        gpu.SaveAsCs16(np.zeros(0, np.float32))                    # "IQ array is empty." (:79-80)
    # round trip at full size on the device: 2^24 samples, |error| <= one CS16 step of the normalised signal
    import torch
    n = 1 << 25
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    gpu.fill_uniform_dev(5, 0, 0, n, x.data_ptr(), 0)
    q = torch.empty(n, dtype=torch.int16, device="cuda")
    m = torch.zeros(1, dtype=torch.float32, device="cuda")
    back = torch.empty_like(x)
    torch.cuda.synchronize()
    lib = gpu._native.lib()
    assert lib.qpsk_cf32_to_cs16_dev(x.data_ptr(), n, q.data_ptr(), m.data_ptr(), None) == 0
    torch.cuda.synchronize()
    mx = float(m.item())
    assert mx == float(x.abs().max().item())
    assert lib.qpsk_cs16_to_cf32_dev(q.data_ptr(), n, mx / 32767.0, back.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert float((back - x).abs().max().item()) <= mx / 32767.0 * 1.0001
