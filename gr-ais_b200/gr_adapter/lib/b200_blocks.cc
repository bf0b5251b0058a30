// b200_blocks.cc -- GNU Radio block adapters over the C-ABI of libb200ais.so.
//
// Each class keeps the reference block's constructor arguments, io signatures, scheduler
// hints and work()/general_work() signature, and delegates the arithmetic to one C-ABI
// call with channels = 1:
//   corr_est_cc_impl            reference lib/corr_est_cc_impl.cc:48-117 (ctor), :132-162, :164-279
//   msk_timing_recovery_cc_impl reference lib/msk_timing_recovery_cc_impl.cc:45-105, :107-206
//   freqest_impl                reference lib/freqest_impl.cc:41-48, :57-88
//   invert_impl                 reference lib/invert_impl.cc:41-68
//   pdu_to_nmea_impl            reference lib/pdu_to_nmea_impl.cc:43-143
// A failing C-ABI call becomes the exception the reference would have thrown
// (std::out_of_range for argument ranges, std::runtime_error otherwise); there is no CPU path.
#include <ais/corr_est_cc.h>
#include <ais/freqest.h>
#include <ais/invert.h>
#include <ais/msk_timing_recovery_cc.h>
#include <ais/pdu_to_nmea.h>

#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>
#endif

#include <b200ais.h>

#include <cmath>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace gr {
namespace ais {

namespace {

void throw_on(int rc)
{
    if (rc == B200AIS_OK)
        return;
    std::string msg = b200ais_last_error();
    if (rc == B200AIS_E_RANGE)
        throw std::out_of_range(msg);
    throw std::runtime_error("b200ais: " + msg);
}

const char *tag_key_name(int key)
{
    switch (key) {
    case B200AIS_TAG_CORR_START:
        return "corr_start";
    case B200AIS_TAG_PHASE_EST:
        return "phase_est";
    case B200AIS_TAG_TIME_EST:
        return "time_est";
    default:
        return "corr_est";
    }
}

} // namespace

// ------------------------------------------------------------------ corr_est_cc

class corr_est_cc_impl : public corr_est_cc
{
public:
    corr_est_cc_impl(const std::vector<gr_complex> &symbols, float sps, unsigned int mark_delay,
                     float threshold)
        : sync_block("corr_est_cc", io_signature::make(1, 1, sizeof(gr_complex)),
                     io_signature::make(1, 2, sizeof(gr_complex))),
          d_src_id(pmt::intern(alias())), d_sps(sps)
    {
        throw_on(b200ais_corr_est_create(&d_h, reinterpret_cast<const float *>(symbols.data()),
                                         (int)symbols.size(), sps, mark_delay, threshold, 1));
        apply_scheduler_hints();
        // one detection per isps items at most, 7 tags each (4 on port 0, 3 debug copies)
        set_max_noutput_items(24 * 1024);
    }
    ~corr_est_cc_impl() override { b200ais_corr_est_destroy(d_h); }

    std::vector<gr_complex> symbols() const override
    {
        int n = 0;
        throw_on(b200ais_corr_est_symbols(d_h, nullptr, 0, &n));
        std::vector<gr_complex> out((size_t)n);
        throw_on(b200ais_corr_est_symbols(d_h, reinterpret_cast<float *>(out.data()), n, &n));
        return out;
    }

    void set_symbols(const std::vector<gr_complex> &symbols) override
    {
        gr::thread::scoped_lock lock(d_setlock);
        throw_on(b200ais_corr_est_set_symbols(d_h, reinterpret_cast<const float *>(symbols.data()),
                                              (int)symbols.size()));
        apply_scheduler_hints();
    }

    int work(int noutput_items, gr_vector_const_void_star &input_items,
             gr_vector_void_star &output_items) override
    {
        gr::thread::scoped_lock lock(d_setlock);
        const float *in = static_cast<const float *>(input_items[0]);
        float *out0 = static_cast<float *>(output_items[0]);
        float *out1 = output_items.size() > 1 ? static_cast<float *>(output_items[1]) : nullptr;
        const int L = (int)history() - 1;
        const int isps = (int)(d_sps + 0.5f);
        const int max_tags = 7 * (noutput_items / (isps > 0 ? isps : 1) + 2);
        if ((int)d_tags.size() < max_tags)
            d_tags.resize((size_t)max_tags);
        int ntags = 0;
        throw_on(b200ais_corr_est_work(d_h, noutput_items, in, (size_t)noutput_items + L,
                                       nitems_written(0), out0, out1, (size_t)noutput_items,
                                       d_tags.data(), max_tags, &ntags));
        for (int k = 0; k < ntags; k++) {
            const b200ais_tag &t = d_tags[(size_t)k];
            add_item_tag((unsigned)t.port, t.offset, pmt::intern(tag_key_name(t.key)),
                         pmt::from_double(t.value), d_src_id);
        }
        return noutput_items;
    }

private:
    void apply_scheduler_hints()
    {
        // reference lib/corr_est_cc_impl.cc:85,95-98: output multiple = fft_filter block size,
        // history = taps + 1, output 0 lags the input by the template length
        set_output_multiple(b200ais_corr_est_output_multiple(d_h));
        set_history((unsigned)b200ais_corr_est_history(d_h));
        declare_sample_delay(1, 0);
        declare_sample_delay(0, (unsigned)b200ais_corr_est_history(d_h) - 1);
    }

    b200ais_corr_est *d_h = nullptr;
    pmt::pmt_t d_src_id;
    float d_sps;
    gr::thread::mutex d_setlock;
    std::vector<b200ais_tag> d_tags;
};

corr_est_cc::sptr corr_est_cc::make(const std::vector<gr_complex> &symbols, float sps,
                                    unsigned int mark_delay, float threshold)
{
    return gnuradio::get_initial_sptr(new corr_est_cc_impl(symbols, sps, mark_delay, threshold));
}

// --------------------------------------------------------- msk_timing_recovery_cc

class msk_timing_recovery_cc_impl : public msk_timing_recovery_cc
{
public:
    msk_timing_recovery_cc_impl(float sps, float gain, float limit, int osps)
        : gr::block("msk_timing_recovery_cc", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                    gr::io_signature::make3(1, 3, sizeof(gr_complex), sizeof(float), sizeof(float))),
          d_osps(osps)
    {
        throw_on(b200ais_msk_create(&d_h, sps, gain, limit, osps, 1));
        set_relative_rate(osps / sps);
        enable_update_rate(true);
    }
    ~msk_timing_recovery_cc_impl() override { b200ais_msk_destroy(d_h); }

    void set_sps(float sps) override
    {
        throw_on(b200ais_msk_set_sps(d_h, sps));
        set_relative_rate(d_osps / sps);
    }
    float get_sps(void) override { return b200ais_msk_get_sps(d_h); }
    void set_gain(float gain) override { throw_on(b200ais_msk_set_gain(d_h, gain)); }
    float get_gain(void) override { return b200ais_msk_get_gain(d_h); }
    void set_limit(float limit) override { throw_on(b200ais_msk_set_limit(d_h, limit)); }
    float get_limit(void) override { return b200ais_msk_get_limit(d_h); }

    void forecast(int noutput_items, gr_vector_int &ninput_items_required) override
    {
        const int need = b200ais_msk_forecast(d_h, noutput_items);
        for (auto &v : ninput_items_required)
            v = need;
    }

    int general_work(int noutput_items, gr_vector_int &ninput_items,
                     gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override
    {
        const float *in = static_cast<const float *>(input_items[0]);
        float *out = static_cast<float *>(output_items[0]);
        float *out2 = output_items.size() >= 2 ? static_cast<float *>(output_items[1]) : nullptr;
        float *out3 = output_items.size() >= 3 ? static_cast<float *>(output_items[2]) : nullptr;
        const int nin = ninput_items[0];

        // every time_est tag on the visible input; the kernel applies the reference's
        // [read, read + ninput - 3*d_sps) window itself
        std::vector<tag_t> tags;
        get_tags_in_range(tags, 0, nitems_read(0), nitems_read(0) + (uint64_t)(nin > 0 ? nin : 0),
                          pmt::intern("time_est"));
        d_tags.resize(tags.size() ? tags.size() : 1);
        for (size_t k = 0; k < tags.size(); k++) {
            d_tags[k].offset = tags[k].offset;
            d_tags[k].key = B200AIS_TAG_TIME_EST;
            d_tags[k].port = 0;
            d_tags[k].value = pmt::to_double(tags[k].value);
        }
        int ntags = (int)tags.size(), nprod = 0, ncons = 0;
        throw_on(b200ais_msk_general_work(d_h, noutput_items, nin, in, (size_t)(nin > 0 ? nin : 1),
                                          nitems_read(0), d_tags.data(), (int)d_tags.size(), &ntags,
                                          out, out2, out3, (size_t)(noutput_items > 0 ? noutput_items : 1),
                                          &nprod, &ncons));
        consume_each(ncons);
        return nprod;
    }

private:
    b200ais_msk *d_h = nullptr;
    int d_osps;
    std::vector<b200ais_tag> d_tags;
};

msk_timing_recovery_cc::sptr msk_timing_recovery_cc::make(float sps, float gain, float limit,
                                                          int osps = 1)
{
    return gnuradio::get_initial_sptr(new msk_timing_recovery_cc_impl(sps, gain, limit, osps));
}

// ---------------------------------------------------------------------- freqest

class freqest_impl : public freqest
{
public:
    freqest_impl(float sample_rate, int data_rate, int fftlen)
        : gr::sync_block("freqest", gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * fftlen),
                         gr::io_signature::make(1, 1, sizeof(float)))
    {
        throw_on(b200ais_freqest_create(&d_h, sample_rate, data_rate, fftlen, 1));
    }
    ~freqest_impl() override { b200ais_freqest_destroy(d_h); }

    int work(int noutput_items, gr_vector_const_void_star &input_items,
             gr_vector_void_star &output_items) override
    {
        throw_on(b200ais_freqest_work(d_h, noutput_items, static_cast<const float *>(input_items[0]),
                                      static_cast<float *>(output_items[0])));
        return noutput_items;
    }

private:
    b200ais_freqest *d_h = nullptr;
};

freqest::sptr freqest::make(float sample_rate, int data_rate, int fftlen)
{
    return gnuradio::get_initial_sptr(new freqest_impl(sample_rate, data_rate, fftlen));
}

// ----------------------------------------------------------------------- invert

class invert_impl : public invert
{
public:
    invert_impl()
        : gr::sync_block("invert", gr::io_signature::make(1, 1, sizeof(char)),
                         gr::io_signature::make(1, 1, sizeof(char)))
    {
    }

    int work(int noutput_items, gr_vector_const_void_star &input_items,
             gr_vector_void_star &output_items) override
    {
        throw_on(b200ais_invert_work(static_cast<const uint8_t *>(input_items[0]),
                                     static_cast<uint8_t *>(output_items[0]), (size_t)noutput_items));
        return noutput_items;
    }
};

invert::sptr invert::make() { return gnuradio::get_initial_sptr(new invert_impl()); }

// ------------------------------------------------------------------ pdu_to_nmea

class pdu_to_nmea_impl : public pdu_to_nmea
{
public:
    explicit pdu_to_nmea_impl(std::string designator)
        : gr::block("pdu_to_nmea", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)),
          d_designator(designator)
    {
        if (designator.empty() || designator.size() > 8)
            throw std::invalid_argument("pdu_to_nmea: designator must be 1..8 characters");
        message_port_register_in(pmt::mp("print"));
        set_msg_handler(pmt::mp("print"), [this](pmt::pmt_t m) { this->print(m); });
        message_port_register_in(pmt::mp("to_nmea"));
        set_msg_handler(pmt::mp("to_nmea"), [this](pmt::pmt_t m) { this->to_nmea(m); });
        message_port_register_out(pmt::mp("out"));
    }

    // msg_to_sentence (reference lib/pdu_to_nmea_impl.cc:127-131) on the device
    std::string msg_to_sentence(pmt::pmt_t msg)
    {
        const uint8_t *p = static_cast<const uint8_t *>(pmt::blob_data(pmt::cdr(msg)));
        const size_t len = pmt::blob_length(pmt::cdr(msg));
        if (len < 1 || len > B200AIS_FRAME_MAX)
            throw std::runtime_error("pdu_to_nmea: PDU length outside [1, 248]");
        b200ais_frame fr;
        memset(&fr, 0, sizeof(fr));
        fr.len = (int32_t)len;
        memcpy(fr.data, p, len);
        const int one = 1;
        char des[8] = {0};
        memcpy(des, d_designator.data(), d_designator.size());
        const int slot = b200ais_nmea_slot_bytes((int)len, d_designator.c_str());
        std::vector<char> out((size_t)slot);
        int n = 0;
        throw_on(b200ais_nmea_format(&fr, &one, 1, 1, des, out.data(), slot, &n));
        if (n < 0)
            throw std::runtime_error("pdu_to_nmea: sentence does not fit");
        return std::string(out.data(), (size_t)n);
    }

    void print(pmt::pmt_t msg) override { std::cout << msg_to_sentence(msg) << std::endl; }

    void to_nmea(pmt::pmt_t msg) override
    {
        std::string aivdm = msg_to_sentence(msg);
        pmt::pmt_t pdu(pmt::cons(pmt::PMT_NIL,
                                 pmt::init_u8vector(aivdm.length(), (const uint8_t *)aivdm.c_str())));
        message_port_pub(pmt::mp("out"), pdu);
    }

    int general_work(int, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &) override
    {
        return 0; // message block: no streams
    }

private:
    std::string d_designator;
};

pdu_to_nmea::sptr pdu_to_nmea::make(std::string designator)
{
    return gnuradio::get_initial_sptr(new pdu_to_nmea_impl(designator));
}

} // namespace ais
} // namespace gr
