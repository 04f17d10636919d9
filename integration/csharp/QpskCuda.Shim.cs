// integration/csharp/QpskCuda.Shim.cs — the reference's public classes on the hot path with their bodies replaced by
// calls into libqpskcuda.so.  Same namespaces, class names, constructor and method signatures, defaults and exceptions as
// Modulation-Simulation/{Models/FIRFilter.cs, Models/RRC-filter.cs, Models/Band-Edge Filter.cs, Models/MuellerMuller.cs,
// Models/CostasLoopQpsk.cs, QPSKModulator.cs, QPSKDeModulator.cs}; each member cites the line it stands in for.
// Loop and filter state lives on the device inside the native handle, one handle per object, exactly as the managed
// objects were one per stream.  HelperFunctions (BitPacker, SaveAsCs16, ...) and RealFIRFilter stay as they are.
// Not built in this repository: the image has no .NET toolchain.
using System;
using System.Text;
using QPSK.Native;

namespace QPSK.Models
{
    public class RRCFilter
    {
        public static unsafe double[] generateCoefficents(double spanSymbols, double beta, int sampleRate, int SymbolRate)   // RRC-filter.cs:16
        {
            QpskCuda.Check(QpskCuda.qpsk_rrc_taps(spanSymbols, beta, sampleRate, SymbolRate, null, 0, out int n));
            var h = new double[n];
            fixed (double* p = h) QpskCuda.Check(QpskCuda.qpsk_rrc_taps(spanSymbols, beta, sampleRate, SymbolRate, p, n, out n));
            return h;
        }
    }

    public sealed unsafe class ComplexFIRFilter : IDisposable
    {
        readonly IntPtr _h;
        public readonly float[] taps;                                                // FIRFilter.cs:11
        public ComplexFIRFilter(float[] tapsInterleavedIQ)                         // :29
        {
            if (tapsInterleavedIQ == null) throw new ArgumentNullException(nameof(tapsInterleavedIQ));
            taps = (float[])tapsInterleavedIQ.Clone();
            fixed (float* t = tapsInterleavedIQ)
                QpskCuda.Check(QpskCuda.qpsk_fir_create(t, tapsInterleavedIQ.Length, out _h), nameof(tapsInterleavedIQ));
        }
        public void Filter(float inI, float inQ, out float outI, out float outQ)   // :59  one sample = a 2-float span
        {
            float* io = stackalloc float[4];
            io[0] = inI; io[1] = inQ;
            QpskCuda.Check(QpskCuda.qpsk_fir_filter(_h, io, io + 2, 2, 2));
            outI = io[2]; outQ = io[3];
        }
        public void Filter(ReadOnlySpan<float> iqIn, Span<float> iqOut)            // :80  streaming, delay line on the device
        {
            fixed (float* i = iqIn) fixed (float* o = iqOut)
                QpskCuda.Check(QpskCuda.qpsk_fir_filter(_h, i, o, iqIn.Length, iqOut.Length), nameof(iqIn));
        }
        public float[] fftFilter(float[] iqData)                                   // :96  stateless, offset N-1
        {
            if (iqData == null) throw new ArgumentNullException(nameof(iqData));
            var y = new float[(iqData.Length & 1) == 0 ? iqData.Length : 0];
            fixed (float* i = iqData) fixed (float* o = y)
                QpskCuda.Check(QpskCuda.qpsk_fir_fft_filter(_h, i, o, iqData.Length), nameof(iqData));
            return y;
        }
        public void Dispose() => QpskCuda.qpsk_fir_destroy(_h);
    }

    public sealed unsafe class FLLBandEdgeFilter : IDisposable
    {
        readonly IntPtr _h;
        public float sps, rolloff, bandwidth;                                        // Band-Edge Filter.cs:19-22
        public int filterSize;
        public float phase                                                           // :25 (a public field upstream)
        {
            get { QpskCuda.Check(QpskCuda.qpsk_fll_get_state(_h, out float p, out _)); return p; }
            set { QpskCuda.Check(QpskCuda.qpsk_fll_get_state(_h, out _, out float f)); QpskCuda.Check(QpskCuda.qpsk_fll_set_state(_h, in value, in f)); }
        }
        public float freq                                                            // :26
        {
            get { QpskCuda.Check(QpskCuda.qpsk_fll_get_state(_h, out _, out float f)); return f; }
            set { QpskCuda.Check(QpskCuda.qpsk_fll_get_state(_h, out float p, out _)); QpskCuda.Check(QpskCuda.qpsk_fll_set_state(_h, in p, in value)); }
        }
        public FLLBandEdgeFilter(float sps, float rolloff, int filterSize, float bandwidth)   // :40 (range checks :42-45 -> status -3)
        {
            this.sps = sps; this.rolloff = rolloff; this.filterSize = filterSize; this.bandwidth = bandwidth;
            QpskCuda.Check(QpskCuda.qpsk_fll_create(sps, rolloff, filterSize, bandwidth, out _h));
        }
        public int Process(ReadOnlySpan<float> inputIQ, Span<float> outputIQ)       // :64
        {
            fixed (float* i = inputIQ) fixed (float* o = outputIQ)
                QpskCuda.Check(QpskCuda.qpsk_fll_process(_h, i, o, inputIQ.Length, outputIQ.Length), nameof(inputIQ));
            return inputIQ.Length;
        }
        public float[] Process(float[] inputIQ)                                      // :90
        {
            if (inputIQ == null) throw new ArgumentNullException(nameof(inputIQ));
            var y = new float[inputIQ.Length];
            Process(inputIQ.AsSpan(), y.AsSpan());
            return y;
        }
        public void Process(float inI, float inQ, out float outI, out float outQ)   // :102
        {
            float* io = stackalloc float[4];
            io[0] = inI; io[1] = inQ;
            QpskCuda.Check(QpskCuda.qpsk_fll_process(_h, io, io + 2, 2, 2));
            outI = io[2]; outQ = io[3];
        }
        public void Dispose() => QpskCuda.qpsk_fll_destroy(_h);
    }

    public sealed unsafe class MuellerMuller : IDisposable
    {
        readonly IntPtr _h;
        public MuellerMuller(double samplesPerSymbol, double kp, double ki)          // MuellerMuller.cs:38
        {
            QpskCuda.Check(QpskCuda.qpsk_mm_create(samplesPerSymbol, kp, ki, out _h));
        }
        public int Process(ReadOnlySpan<float> incomingMfSamplesIQ, Span<float> outputSymbolsIQ)   // :52
        {
            int nSym;
            fixed (float* i = incomingMfSamplesIQ) fixed (float* o = outputSymbolsIQ)
                QpskCuda.Check(QpskCuda.qpsk_mm_process(_h, i, incomingMfSamplesIQ.Length, o, outputSymbolsIQ.Length, out nSym),
                               nameof(incomingMfSamplesIQ));
            return 2 * nSym;      // the native call counts symbols, the reference returns floats written (:135)
        }
        public float[] Process(float[] incomingMfSamplesIQ)                          // :141
        {
            if (incomingMfSamplesIQ == null) throw new ArgumentNullException(nameof(incomingMfSamplesIQ));
            if ((incomingMfSamplesIQ.Length & 1) != 0)
                throw new ArgumentException("Input must be interleaved IQ with even length.", nameof(incomingMfSamplesIQ));
            var tmp = new float[incomingMfSamplesIQ.Length];
            int n = Process(incomingMfSamplesIQ.AsSpan(), tmp.AsSpan());
            var y = new float[n];
            Array.Copy(tmp, y, n);
            return y;
        }
        public void Dispose() => QpskCuda.qpsk_mm_destroy(_h);
    }

    public sealed unsafe class CostasLoopQpsk : IDisposable
    {
        readonly IntPtr _h;
        public CostasLoopQpsk(double sampleRate, double loopBandwidthHz, double damping = 0.707)   // CostasLoopQpsk.cs:29
        {
            QpskCuda.Check(QpskCuda.qpsk_costas_create(sampleRate, loopBandwidthHz, damping, out _h));
        }
        public static void GetSign(float i, float q, out float di, out float dq)    // :52 (pure host helper, unchanged)
        {
            di = (i >= 0f) ? 1f : -1f;
            dq = (q >= 0f) ? 1f : -1f;
        }
        public void Process(float inI, float inQ, out float outI, out float outQ)   // :63
        {
            float* io = stackalloc float[4];
            io[0] = inI; io[1] = inQ;
            QpskCuda.Check(QpskCuda.qpsk_costas_process(_h, io, io + 2, 2, 2));
            outI = io[2]; outQ = io[3];
        }
        public int Process(ReadOnlySpan<float> iqIn, Span<float> iqOut)             // :98
        {
            fixed (float* i = iqIn) fixed (float* o = iqOut)
                QpskCuda.Check(QpskCuda.qpsk_costas_process(_h, i, o, iqIn.Length, iqOut.Length), nameof(iqIn));
            return iqIn.Length;
        }
        public float[] Process(float[] iqIn)                                         // :119
        {
            if (iqIn == null) throw new ArgumentNullException(nameof(iqIn));
            var y = new float[iqIn.Length];
            Process(iqIn.AsSpan(), y.AsSpan());
            return y;
        }
        public (double theta, double freq) GetState()                                // :130
        {
            QpskCuda.Check(QpskCuda.qpsk_costas_get_state(_h, out double t, out double f));
            return (t, f);
        }
        public void Dispose() => QpskCuda.qpsk_costas_destroy(_h);
    }
}

namespace QPSK
{
    public sealed unsafe class QPSKModulator : IDisposable
    {
        readonly IntPtr _h;
        public long baudRate;                                                        // QPSKModulator.cs:34
        public QPSKModulator(int SampleRate, int SymbolRate, double RrcAlpha = 0.9, int rrcSpan = 6,
                             bool differentialEncoding = true, string? tsc = null)   // :18
        {
            baudRate = 2L * SymbolRate / 8L;
            QpskCuda.Check(QpskCuda.qpsk_mod_create(SampleRate, SymbolRate, RrcAlpha, rrcSpan, differentialEncoding ? 1 : 0, tsc, out _h));
        }
        public double[] getCoeef()                                                   // :32
        {
            QpskCuda.Check(QpskCuda.qpsk_mod_taps(_h, null, 0, out int n));
            var h = new double[n];
            fixed (double* p = h) QpskCuda.Check(QpskCuda.qpsk_mod_taps(_h, p, n, out n));
            return h;
        }
        public float[] ModulateBytes(ReadOnlySpan<byte> payload, ReadOnlySpan<byte> startMarker, ReadOnlySpan<byte> endMarker,
                                     bool pulseShaping = true)                      // :54  framing START|payload|END on the device
        {
            int ps = pulseShaping ? 1 : 0;
            fixed (byte* p = payload) fixed (byte* s = startMarker) fixed (byte* e = endMarker)
            {
                QpskCuda.Check(QpskCuda.qpsk_mod_modulate_bytes(_h, p, payload.Length, s, startMarker.Length, e, endMarker.Length, ps,
                                                                null, 0, out long n));
                var y = new float[n];
                if (n > 0)
                    fixed (float* o = y)
                        QpskCuda.Check(QpskCuda.qpsk_mod_modulate_bytes(_h, p, payload.Length, s, startMarker.Length, e, endMarker.Length,
                                                                        ps, o, n, out n));
                return y;
            }
        }
        public float[] ModulateTextUtf8(string text, string startMarker = "\u0002", string endMarker = "\u0003",
                                        bool pulseShaping = true, Encoding? encoding = null)   // :74  stays managed + ModulateBytes
        {
            encoding ??= Encoding.UTF8;
            return ModulateBytes(encoding.GetBytes(text ?? string.Empty), encoding.GetBytes(startMarker ?? string.Empty),
                                 encoding.GetBytes(endMarker ?? string.Empty), pulseShaping);
        }
        public float[] Modulate(string data, bool pulseShaping = true)              // :104
        {
            if (data == null) throw new ArgumentNullException(nameof(data));
            byte[] bits = Encoding.Latin1.GetBytes(data);                            // one char per bit, as upstream
            int ps = pulseShaping ? 1 : 0;
            fixed (byte* b = bits)
            {
                QpskCuda.Check(QpskCuda.qpsk_mod_modulate_bits(_h, b, bits.Length, ps, null, 0, out long n));
                var y = new float[n];
                if (n > 0)
                    fixed (float* o = y)
                        QpskCuda.Check(QpskCuda.qpsk_mod_modulate_bits(_h, b, bits.Length, ps, o, n, out n));
                return y;
            }
        }
        public void Dispose() => QpskCuda.qpsk_mod_destroy(_h);
    }

    public sealed unsafe class QPSKDeModulator : IDisposable
    {
        readonly IntPtr _h;
        public QPSKDeModulator(int SampleRate, int SymbolRate, float RrcAlpha = 0.9f, int rrcSpan = 6,
                               double SymbolSyncBandwith = 0.0001, double CostasLoopBandwith = 120, double CFOLoopBandwith = 0.0001f,
                               bool differentialEncoding = true, string? tsc = null)   // QPSKDeModulator.cs:11
        {
            // use_fll = 0: the fll.Process call is commented out upstream (:359, :435); max_frame_bytes = 0: library default
            QpskCuda.Check(QpskCuda.qpsk_demod_create(SampleRate, SymbolRate, RrcAlpha, rrcSpan, SymbolSyncBandwith, CostasLoopBandwith,
                                                      CFOLoopBandwith, differentialEncoding ? 1 : 0, tsc, 0, 0, out _h));
        }
        public byte[] DeModulateBytes(ReadOnlySpan<float> samplesIQ, ReadOnlySpan<byte> startMarker, ReadOnlySpan<byte> endMarker)   // :169
        {
            if (startMarker.Length == 0) throw new ArgumentException("startMarker cannot be empty.", nameof(startMarker));
            if (endMarker.Length == 0) throw new ArgumentException("endMarker cannot be empty.", nameof(endMarker));
            // a frame may have been accumulating on the device over earlier calls: size for the library's frame bound (1 MiB)
            var buf = new byte[Math.Max(samplesIQ.Length / 8 + 64, 1 << 20)];
            fixed (float* i = samplesIQ) fixed (byte* s = startMarker) fixed (byte* e = endMarker) fixed (byte* o = buf)
            {
                QpskCuda.Check(QpskCuda.qpsk_demod_bytes(_h, i, samplesIQ.Length, s, startMarker.Length, e, endMarker.Length, o, buf.Length,
                                                         out long n), nameof(samplesIQ));
                var payload = new byte[n];
                Array.Copy(buf, payload, n);
                return payload;
            }
        }
        public string DeModulateTextUtf8(ReadOnlySpan<float> samplesIQ, string startMarker = "\u0002", string endMarker = "\u0003",
                                         Encoding? encoding = null)                 // :262
        {
            encoding ??= Encoding.UTF8;
            byte[] p = DeModulateBytes(samplesIQ, encoding.GetBytes(startMarker), encoding.GetBytes(endMarker));
            return p.Length == 0 ? string.Empty : encoding.GetString(p);
        }
        public string DeModulate(float[] SamplesIQ) => DeModulate(SamplesIQ.AsSpan());   // :339
        public string DeModulate(ReadOnlySpan<float> SamplesIQ)                      // :345  '0'/'1' chars, TSC already stripped
        {
            var buf = new byte[Math.Max(SamplesIQ.Length, 16)];
            fixed (float* i = SamplesIQ) fixed (byte* o = buf)
            {
                QpskCuda.Check(QpskCuda.qpsk_demod_bits(_h, i, SamplesIQ.Length, o, buf.Length, out long n), nameof(SamplesIQ));
                return Encoding.ASCII.GetString(buf, 0, (int)n);
            }
        }
        public float[] deModulateConstellation(ReadOnlySpan<float> SamplesIQ)        // :427
        {
            var buf = new float[Math.Max(SamplesIQ.Length, 2)];
            fixed (float* i = SamplesIQ) fixed (float* o = buf)
            {
                QpskCuda.Check(QpskCuda.qpsk_demod_constellation(_h, i, SamplesIQ.Length, o, buf.Length, out long nSym), nameof(SamplesIQ));
                var y = new float[2 * nSym];
                Array.Copy(buf, y, 2 * nSym);
                return y;
            }
        }
        public void Dispose() => QpskCuda.qpsk_demod_destroy(_h);
    }
}
