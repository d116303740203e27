/*
 * Headless JUCE shim -- TEST INFRASTRUCTURE ONLY (part of oracle/).
 *
 * The reference (SeanSoraghan/Feature-Extractor) builds against JUCE 4.2.3
 * (Feature-Extractor.jucer:5, module path ../JUCE/modules), which is neither vendored
 * in /root/reference nor available offline.  This header supplies just enough of the
 * JUCE surface for the six hot-path headers (AudioDataCollector.h, RealTimeAudioAnalysis.h,
 * PitchAnalyser.h, SpectralCharacteristics.h, HarmonicCharacteristics.h, RealTimeAnalyser.h)
 * to compile unmodified, headless, with g++.
 *
 * The arithmetic pieces restate the published behaviour of JUCE 4.2.3's
 * juce_audio_basics module (AudioSampleBuffer, FFT) in our own code:
 *   - FFT: un-normalised forward e^{-2 pi i k n / N}, fp32, twiddles evaluated in double and
 *     rounded to float, mixed radix-4-then-radix-2 decimation in time (kissfft lineage);
 *     real-only forward = full complex transform of (x, 0); real-only inverse = full complex
 *     inverse scaled by 1/N and de-interleaved (d[i] = Re, d[i + N] = Im).
 *   - AudioSampleBuffer::applyGainRamp: additive fp32 gain accumulation.
 *   - AudioSampleBuffer::getRMSLevel: fp32 squares accumulated in double.
 *   - the "isClear" short-cuts of AudioBuffer.
 * Parity is therefore pinned to THIS restatement of JUCE, not to a JUCE binary
 * ("parity unpinned" at the JUCE boundary).
 *
 * <math.h> and <stdlib.h> MUST come first: the reference calls unqualified abs(double),
 * log10(float), exp(float).  Its real toolchains (MSVC 2015 / Xcode) pick the floating
 * overloads; with only <cmath>/<cstdlib> libstdc++ would bind abs(double) to int abs(int).
 */
#ifndef FX_ORACLE_JUCE_SHIM_H
#define FX_ORACLE_JUCE_SHIM_H

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <string>
#include <functional>
#include <sstream>
#include <iostream>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// macros / small helpers
#define jassert(expr)      ((void) 0)
#define jassertfalse       ((void) 0)
#define DBG(x)             ((void) 0)
#define JUCE_COMPILER_WARNING(x)
#define JUCE_LIVE_CONSTANT(x) (x)
#define JUCE_DECLARE_NON_COPYABLE(ClassName) \
    ClassName (const ClassName&) = delete;   \
    ClassName& operator= (const ClassName&) = delete;
#define JUCE_DECLARE_NON_COPYABLE_WITH_LEAK_DETECTOR(ClassName) JUCE_DECLARE_NON_COPYABLE (ClassName)

const double double_Pi = 3.1415926535897932384626433832795;
const float  float_Pi  = 3.14159265358979323846f;

template <typename... Ts> void ignoreUnused (const Ts&...) noexcept {}

template <typename T> T jmax (T a, T b)           { return a < b ? b : a; }
template <typename T> T jmax (T a, T b, T c)      { return jmax (a, jmax (b, c)); }
template <typename T> T jmax (T a, T b, T c, T d) { return jmax (a, jmax (b, c, d)); }
template <typename T> T jmin (T a, T b)           { return b < a ? b : a; }

template <typename T, size_t n> int numElementsInArray (T (&)[n]) { return (int) n; }

// ---------------------------------------------------------------------------------------------
// String: only what the debug printers in the hot-path headers touch.
class String
{
public:
    String() {}
    String (const char* t) : s (t) {}
    String (const std::string& t) : s (t) {}
    explicit String (float v)  { std::ostringstream o; o << v; s = o.str(); }
    explicit String (double v) { std::ostringstream o; o << v; s = o.str(); }
    explicit String (int v)    { std::ostringstream o; o << v; s = o.str(); }

    String operator+ (const String& o) const { return String (s + o.s); }
    String operator+ (const char* o)   const { return String (s + o); }
    String& operator<< (const String& o)     { s += o.s; return *this; }
    String& operator<< (const char* o)       { s += o;   return *this; }
    bool operator== (const String& o) const  { return s == o.s; }
    bool operator!= (const String& o) const  { return s != o.s; }

    static const String empty;
    std::string s;
};
const String String::empty;

// ---------------------------------------------------------------------------------------------
template <typename T>
struct Atomic
{
    Atomic() : value (0) {}
    Atomic (T v) : value (v) {}
    T get() const   { return value; }
    void set (T v)  { value = v; }
    volatile T value;
};

template <typename T>
class Point
{
public:
    Point() : x (0), y (0) {}
    Point (T xx, T yy) : x (xx), y (yy) {}
    T getX() const { return x; }
    T getY() const { return y; }
    T x, y;
};

template <typename T>
class Range
{
public:
    Range() : start (0), end (0) {}
    Range (T s, T e) : start (s), end (e) {}
    T getStart() const { return start; }
    T getEnd()   const { return end; }
    void setStart (T s) { start = s; if (end < s) end = s; }
    void setEnd (T e)   { end = e; if (e < start) start = e; }
    T start, end;
};

// ---------------------------------------------------------------------------------------------
// ReferenceCountedObject / ReferenceCountedObjectPtr / HeapBlock: only what the legacy offline analyser's headers
// (AudioFeatures.h, AudioAnalysis.h) need to compile; the drivers never share ownership
class ReferenceCountedObject
{
public:
    void incReferenceCount() noexcept { ++refCount; }
    void decReferenceCount() noexcept { if (--refCount == 0) delete this; }
    int getReferenceCount() const noexcept { return refCount; }
protected:
    ReferenceCountedObject() : refCount (0) {}
    virtual ~ReferenceCountedObject() {}
private:
    int refCount;
};

template <class T>
class ReferenceCountedObjectPtr
{
public:
    ReferenceCountedObjectPtr() : obj (nullptr) {}
    ReferenceCountedObjectPtr (T* o) : obj (o) { if (obj) obj->incReferenceCount(); }
    ReferenceCountedObjectPtr (const ReferenceCountedObjectPtr& o) : obj (o.obj) { if (obj) obj->incReferenceCount(); }
    ~ReferenceCountedObjectPtr() { if (obj) obj->decReferenceCount(); }
    ReferenceCountedObjectPtr& operator= (const ReferenceCountedObjectPtr& o)
    {
        if (o.obj) o.obj->incReferenceCount();
        if (obj) obj->decReferenceCount();
        obj = o.obj;
        return *this;
    }
    T* operator->() const { return obj; }
    T* get() const { return obj; }
    operator T*() const { return obj; }
private:
    T* obj;
};

template <class T>
class HeapBlock
{
public:
    HeapBlock() {}
    explicit HeapBlock (size_t n) : data (n) {}
    void allocate (size_t n, bool) { data.assign (n, T()); }
    T* getData() { return data.data(); }
    T& operator[] (size_t i) { return data[i]; }
    operator T*() { return data.data(); }
private:
    std::vector<T> data;
};

// ---------------------------------------------------------------------------------------------
// AudioSampleBuffer  (juce::AudioBuffer<float>, JUCE 4.2.x semantics incl. the isClear flag)
class AudioSampleBuffer
{
public:
    AudioSampleBuffer() : numChannels (0), size (0), isClear (false) {}

    AudioSampleBuffer (int numChannelsToAllocate, int numSamplesToAllocate)
        : numChannels (numChannelsToAllocate), size (numSamplesToAllocate), isClear (false)
    {
        // JUCE leaves fresh storage uninitialised; zero it so the oracle is deterministic.
        data.assign ((size_t) numChannels, std::vector<float> ((size_t) size, 0.0f));
    }

    AudioSampleBuffer (const AudioSampleBuffer& other)
        : numChannels (other.numChannels), size (other.size), isClear (false)
    {
        data.assign ((size_t) numChannels, std::vector<float> ((size_t) size, 0.0f));
        if (other.isClear) clear();
        else               data = other.data;
    }

    AudioSampleBuffer& operator= (const AudioSampleBuffer& other)
    {
        if (this != &other)
        {
            setSize (other.numChannels, other.size, false, false, false);
            if (other.isClear) clear();
            else             { isClear = false; data = other.data; }
        }
        return *this;
    }

    int getNumChannels() const noexcept { return numChannels; }
    int getNumSamples()  const noexcept { return size; }

    const float* getReadPointer (int channel) const noexcept              { return data[(size_t) channel].data(); }
    const float* getReadPointer (int channel, int sampleIndex) const noexcept { return data[(size_t) channel].data() + sampleIndex; }
    float* getWritePointer (int channel) noexcept                         { isClear = false; return data[(size_t) channel].data(); }
    float* getWritePointer (int channel, int sampleIndex) noexcept        { isClear = false; return data[(size_t) channel].data() + sampleIndex; }

    float getSample (int channel, int sampleIndex) const noexcept         { return data[(size_t) channel][(size_t) sampleIndex]; }
    void  setSample (int channel, int destSample, float newValue) noexcept
    {
        data[(size_t) channel][(size_t) destSample] = newValue;
        isClear = false;
    }

    void setSize (int newNumChannels, int newNumSamples,
                  bool keepExistingContent = false, bool clearExtraSpace = false, bool avoidReallocating = false)
    {
        ignoreUnused (avoidReallocating);
        if (newNumSamples == size && newNumChannels == numChannels)
            return;

        std::vector<std::vector<float>> fresh ((size_t) newNumChannels, std::vector<float> ((size_t) newNumSamples, 0.0f));
        if (keepExistingContent)
        {
            const int chans = jmin (newNumChannels, numChannels);
            const int n     = jmin (newNumSamples, size);
            for (int c = 0; c < chans; ++c)
                std::copy (data[(size_t) c].begin(), data[(size_t) c].begin() + n, fresh[(size_t) c].begin());
        }
        // (clearExtraSpace || isClear) -> zeroed; otherwise JUCE leaves garbage, we leave zeros.
        ignoreUnused (clearExtraSpace);
        data.swap (fresh);
        numChannels = newNumChannels;
        size = newNumSamples;
    }

    void clear() noexcept
    {
        if (! isClear)
        {
            for (auto& ch : data) std::fill (ch.begin(), ch.end(), 0.0f);
            isClear = true;
        }
    }

    void applyGain (int startSample, int numSamples, float gain) noexcept
    {
        for (int ch = 0; ch < numChannels; ++ch) applyGain (ch, startSample, numSamples, gain);
    }
    void applyGain (int channel, int startSample, int numSamples, float gain) noexcept
    {
        if (gain != 1.0f && ! isClear)
        {
            float* d = data[(size_t) channel].data() + startSample;
            if (gain == 0.0f) for (int i = 0; i < numSamples; ++i) d[i] = 0.0f;
            else              for (int i = 0; i < numSamples; ++i) d[i] *= gain;
        }
    }

    void applyGainRamp (int channel, int startSample, int numSamples, float startGain, float endGain) noexcept
    {
        if (! isClear)
        {
            if (startGain == endGain)
            {
                applyGain (channel, startSample, numSamples, startGain);
            }
            else
            {
                const float increment = (endGain - startGain) / numSamples;
                float* d = data[(size_t) channel].data() + startSample;
                while (--numSamples >= 0)
                {
                    *d++ *= startGain;
                    startGain += increment;
                }
            }
        }
    }

    void copyFrom (int destChannel, int destStartSample, const AudioSampleBuffer& source,
                   int sourceChannel, int sourceStartSample, int numSamples) noexcept
    {
        if (numSamples > 0)
        {
            if (source.isClear)
            {
                if (! isClear)
                    std::fill_n (data[(size_t) destChannel].begin() + destStartSample, numSamples, 0.0f);
            }
            else
            {
                isClear = false;
                std::copy_n (source.data[(size_t) sourceChannel].begin() + sourceStartSample, numSamples,
                             data[(size_t) destChannel].begin() + destStartSample);
            }
        }
    }

    void copyFrom (int destChannel, int destStartSample, const float* source, int numSamples) noexcept
    {
        if (numSamples > 0)
        {
            isClear = false;
            std::copy_n (source, numSamples, data[(size_t) destChannel].begin() + destStartSample);
        }
    }

    void copyFrom (int destChannel, int destStartSample, const float* source, int numSamples, float gain) noexcept
    {
        if (numSamples > 0)
        {
            float* d = data[(size_t) destChannel].data() + destStartSample;
            if (gain != 1.0f)
            {
                if (gain == 0.0f)
                {
                    if (! isClear) std::fill_n (d, numSamples, 0.0f);
                }
                else
                {
                    isClear = false;
                    for (int i = 0; i < numSamples; ++i) d[i] = source[i] * gain;
                }
            }
            else
            {
                isClear = false;
                std::copy_n (source, numSamples, d);
            }
        }
    }

    Range<float> findMinMax (int channel, int startSample, int numSamples) const noexcept
    {
        if (isClear || numSamples <= 0)
            return Range<float>();
        const float* d = data[(size_t) channel].data() + startSample;
        float mn = d[0], mx = d[0];
        for (int i = 1; i < numSamples; ++i)
        {
            if (d[i] < mn) mn = d[i];
            if (d[i] > mx) mx = d[i];
        }
        return Range<float> (mn, mx);
    }

    float getMagnitude (int channel, int startSample, int numSamples) const noexcept
    {
        if (isClear)
            return 0.0f;
        const Range<float> r (findMinMax (channel, startSample, numSamples));
        return jmax (r.getStart(), -r.getStart(), r.getEnd(), -r.getEnd());
    }

    float getRMSLevel (int channel, int startSample, int numSamples) const noexcept
    {
        if (numSamples <= 0 || channel < 0 || channel >= numChannels || isClear)
            return 0.0f;
        const float* d = data[(size_t) channel].data() + startSample;
        double sum = 0.0;
        for (int i = 0; i < numSamples; ++i)
        {
            const float sample = d[i];
            sum += sample * sample;          // fp32 square, fp64 accumulate
        }
        return (float) sqrt (sum / numSamples);
    }

private:
    int numChannels, size;
    std::vector<std::vector<float>> data;
    bool isClear;
};

// ---------------------------------------------------------------------------------------------
// FFT  (juce::FFT of juce_audio_basics 4.2.x, restated)
class FFT
{
public:
    struct Complex { float r, i; };

    FFT (int order, bool isInverse) : size (1 << order), inverse (isInverse), twiddle ((size_t) size)
    {
        // juce_FFT.cpp (4.2.x) [JUCE-recall]: inverseFactor = (isInverse ? 2.0 : -2.0) * double_Pi / fftSize; phase = i * inverseFactor
        // (for the power-of-two sizes used here the scaling by 1 / size is exact, so the order of operations cannot change a bit)
        const double inverseFactor = (inverse ? 2.0 : -2.0) * double_Pi / size;
        for (int i = 0; i < size; ++i)
        {
            const double phase = i * inverseFactor;
            twiddle[(size_t) i].r = (float) cos (phase);
            twiddle[(size_t) i].i = (float) sin (phase);
        }
        // radix plan: as many 4s as divide, then 2s (sizes are powers of two here)
        int n = size;
        while (n > 1)
        {
            const int radix = (n % 4 == 0) ? 4 : 2;
            n /= radix;
            Stage st; st.radix = radix; st.length = n;
            plan.push_back (st);
        }
        if (plan.empty()) { Stage st; st.radix = 1; st.length = 1; plan.push_back (st); }
    }

    int getSize() const noexcept { return size; }

    void perform (const Complex* input, Complex* output) const noexcept
    {
        recurse (input, output, 1, 0);
    }

    void performRealOnlyForwardTransform (float* d) const noexcept
    {
        std::vector<Complex> scratch ((size_t) size);
        for (int i = 0; i < size; ++i) { scratch[(size_t) i].r = d[i]; scratch[(size_t) i].i = 0.0f; }
        perform (scratch.data(), reinterpret_cast<Complex*> (d));
    }

    void performRealOnlyInverseTransform (float* d) const noexcept
    {
        std::vector<Complex> scratch ((size_t) size);
        perform (reinterpret_cast<const Complex*> (d), scratch.data());
        const float scaleFactor = 1.0f / size;
        for (int i = 0; i < size; ++i)
        {
            d[i]        = scratch[(size_t) i].r * scaleFactor;
            d[i + size] = scratch[(size_t) i].i * scaleFactor;
        }
    }

    void performFrequencyOnlyForwardTransform (float* d) const noexcept
    {
        performRealOnlyForwardTransform (d);
        const int twiceSize = size * 2;
        for (int i = 0; i < twiceSize; i += 2)
        {
            d[i / 2] = (float) sqrt ((double) d[i] * d[i] + (double) d[i + 1] * d[i + 1]);
            if (i >= size) { d[i] = 0; d[i + 1] = 0; }
        }
    }

private:
    struct Stage { int radix, length; };

    static Complex cmul (Complex a, Complex b) noexcept
    {
        Complex c = { a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r };
        return c;
    }
    static Complex cadd (Complex a, Complex b) noexcept { Complex c = { a.r + b.r, a.i + b.i }; return c; }
    static Complex csub (Complex a, Complex b) noexcept { Complex c = { a.r - b.r, a.i - b.i }; return c; }

    // decimation in time: split the input comb of this level into `radix` interleaved combs,
    // transform each into a contiguous block of `length` outputs, then combine in place.
    void recurse (const Complex* in, Complex* out, int stride, size_t level) const noexcept
    {
        const Stage st = plan[level];
        if (st.radix == 1) { *out = *in; return; }

        if (st.length == 1)
        {
            for (int q = 0; q < st.radix; ++q)
                out[q] = in[q * stride];
        }
        else
        {
            for (int q = 0; q < st.radix; ++q)
                recurse (in + q * stride, out + q * st.length, stride * st.radix, level + 1);
        }

        if (st.radix == 4) combine4 (out, stride, st.length);
        else               combine2 (out, stride, st.length);
    }

    void combine2 (Complex* d, int stride, int length) const noexcept
    {
        Complex* hi = d + length;
        const Complex* tw = twiddle.data();
        for (int i = 0; i < length; ++i)
        {
            const Complex s = cmul (*hi, *tw);
            tw += stride;
            *hi = csub (*d, s);
            *d  = cadd (*d, s);
            ++hi; ++d;
        }
    }

    void combine4 (Complex* d, int stride, int length) const noexcept
    {
        const int l2 = length * 2, l3 = length * 3;
        const Complex* t1 = twiddle.data();
        const Complex* t2 = t1;
        const Complex* t3 = t1;
        for (int i = 0; i < length; ++i)
        {
            const Complex s0 = cmul (d[length], *t1);
            const Complex s1 = cmul (d[l2],     *t2);
            const Complex s2 = cmul (d[l3],     *t3);
            const Complex s3 = cadd (s0, s2);
            const Complex s4 = csub (s0, s2);
            const Complex s5 = csub (*d, s1);
            *d = cadd (*d, s1);
            d[l2] = csub (*d, s3);
            t1 += stride; t2 += stride * 2; t3 += stride * 3;
            *d = cadd (*d, s3);
            if (inverse)
            {
                d[length].r = s5.r - s4.i;  d[length].i = s5.i + s4.r;
                d[l3].r     = s5.r + s4.i;  d[l3].i     = s5.i - s4.r;
            }
            else
            {
                d[length].r = s5.r + s4.i;  d[length].i = s5.i - s4.r;
                d[l3].r     = s5.r - s4.i;  d[l3].i     = s5.i + s4.r;
            }
            ++d;
        }
    }

    int size;
    bool inverse;
    std::vector<Complex> twiddle;
    std::vector<Stage> plan;
};

// ---------------------------------------------------------------------------------------------
// Thread: headless, single-stepped.  run() bodies in the reference are
//   while (! threadShouldExit()) { ...one hop... ; wait (-1); }
// stepOnce() lets exactly one iteration execute on the calling thread.
class Thread
{
public:
    explicit Thread (const String& name) : threadName (name), iterationsLeft (0) {}
    virtual ~Thread() {}
    virtual void run() = 0;

    bool threadShouldExit() const      { return iterationsLeft-- <= 0; }
    bool wait (int) const              { return true; }
    void notify() const                {}
    void startThread (int = 5)         {}
    bool stopThread (int)              { return true; }
    void stepOnce()                    { iterationsLeft = 1; run(); }

private:
    String threadName;
    mutable int iterationsLeft;
};

class AudioIODevice;

class AudioIODeviceCallback
{
public:
    virtual ~AudioIODeviceCallback() {}
    virtual void audioDeviceIOCallback (const float** inputChannelData, int numInputChannels,
                                        float** outputChannelData, int numOutputChannels, int numSamples) = 0;
    virtual void audioDeviceAboutToStart (AudioIODevice* device) = 0;
    virtual void audioDeviceStopped() = 0;
};

#endif
