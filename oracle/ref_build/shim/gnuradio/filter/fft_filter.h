// Shim of <gnuradio/filter/fft_filter.h>: filter::kernel::fft_filter_ccc with GNU Radio 3.8's
// interface (ctor(decimation, taps, nthreads), set_taps() -> nsamples, filter(nitems, in, out)),
// as corr_est_cc_impl.cc:77,84,146,188 uses it.  The arithmetic [G] is the oracle's
// restatement (ao_fftfilt_*).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <gnuradio/block.h>

#include "ais_oracle.h"

namespace gr {
namespace filter {
namespace kernel {
class fft_filter_ccc
{
public:
    fft_filter_ccc(int decimation, const std::vector<gr_complex> &taps, int nthreads = 1)
    {
        (void)nthreads;
        if (decimation != 1)
            throw std::invalid_argument("fft_filter_ccc shim: decimation 1 only");
        ao_fftfilt_init(&d_f);
        set_taps(taps);
    }
    ~fft_filter_ccc() { ao_fftfilt_free(&d_f); }
    fft_filter_ccc(const fft_filter_ccc &) = delete;
    fft_filter_ccc &operator=(const fft_filter_ccc &) = delete;
    int set_taps(const std::vector<gr_complex> &taps)
    {
        d_taps = taps;
        return ao_fftfilt_set_taps(&d_f, reinterpret_cast<const float *>(taps.data()), (int)taps.size());
    }
    std::vector<gr_complex> taps() const { return d_taps; }
    unsigned int ntaps() const { return (unsigned)d_taps.size(); }
    int filter(int nitems, const gr_complex *input, gr_complex *output)
    {
        return ao_fftfilt_filter(&d_f, nitems, reinterpret_cast<const float *>(input),
                                 reinterpret_cast<float *>(output));
    }

private:
    ao_fftfilt d_f;
    std::vector<gr_complex> d_taps;
};
} // namespace kernel
} // namespace filter
} // namespace gr
