// ref_harness.cc -- C entry points that drive the gr-ais reference's OWN block classes
// (/root/reference/lib/*_impl.cc, compiled unmodified by oracle/ref_build/Makefile) the way the
// GNU Radio scheduler would: item counters, history, stream tags in and out.
//
// TEST INFRASTRUCTURE ONLY: loaded by tests/ (to pin oracle/ais_oracle.c and the CUDA path to
// the reference's code) and by bench.py's CPU arm.  The product never links or loads it.
//
// The handles are exported twice: as plain functions (ref_*) for ctypes, and as an ao_blocks
// table (ref_blocks()) that plugs the reference's classes into the oracle's chain/stream
// schedule in place of the restated blocks.
#include <ais/corr_est_cc.h>
#include <ais/freqest.h>
#include <ais/invert.h>
#include <ais/msk_timing_recovery_cc.h>
#include <ais/pdu_to_nmea.h>

#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "ais_oracle.h"

namespace {

std::vector<gr_complex> to_vec(const float *iq, int n)
{
    const gr_complex *p = reinterpret_cast<const gr_complex *>(iq);
    return std::vector<gr_complex>(p, p + n);
}

int tag_key(const pmt::pmt_t &k)
{
    const std::string s = pmt::symbol_to_string(k);
    if (s == "corr_start")
        return AO_TAG_CORR_START;
    if (s == "phase_est")
        return AO_TAG_PHASE_EST;
    if (s == "time_est")
        return AO_TAG_TIME_EST;
    if (s == "corr_est")
        return AO_TAG_CORR_EST;
    return -1;
}

const char *key_name(int k)
{
    static const char *names[] = { "corr_start", "phase_est", "time_est", "corr_est" };
    return (k >= 0 && k < 4) ? names[k] : "unknown";
}

struct RefMsk {
    gr::ais::msk_timing_recovery_cc::sptr blk;
    gr_complex prev = gr_complex(0, 0); // in[-1]: the last item consumed by the previous call
    std::vector<gr_complex> buf;
};

} // namespace

extern "C" {

// ---------------------------------------------------------------- corr_est_cc
void *ref_corr_est_new(const float *symbols, int L, float sps, unsigned mark_delay, float threshold)
{
    try {
        return new gr::ais::corr_est_cc::sptr(
            gr::ais::corr_est_cc::make(to_vec(symbols, L), sps, mark_delay, threshold));
    } catch (...) {
        return 0;
    }
}

void ref_corr_est_delete(void *h) { delete static_cast<gr::ais::corr_est_cc::sptr *>(h); }

// scheduler hints the constructor / set_symbols left on the block
// (corr_est_cc_impl.cc:85,95-98,112): history, output multiple, max noutput, sample delays
void ref_corr_est_hints(void *h, int *history, int *output_multiple, int *max_noutput, int *delay0,
                        int *delay1)
{
    gr::ais::corr_est_cc &b = **static_cast<gr::ais::corr_est_cc::sptr *>(h);
    *history = (int)b.history();
    *output_multiple = b.output_multiple();
    *max_noutput = b.max_noutput_items();
    *delay0 = (int)b.sample_delay(0);
    *delay1 = (int)b.sample_delay(1);
}

int ref_corr_est_output_multiple(void *h)
{
    return (*static_cast<gr::ais::corr_est_cc::sptr *>(h))->output_multiple();
}

int ref_corr_est_symbols(void *h, float *out, int cap)
{
    std::vector<gr_complex> s = (*static_cast<gr::ais::corr_est_cc::sptr *>(h))->symbols();
    int n = (int)s.size() < cap ? (int)s.size() : cap;
    memcpy(out, s.data(), sizeof(gr_complex) * (size_t)n);
    return (int)s.size();
}

int ref_corr_est_set_symbols(void *h, const float *symbols, int L)
{
    (*static_cast<gr::ais::corr_est_cc::sptr *>(h))->set_symbols(to_vec(symbols, L));
    return 0;
}

// One work() call.  in holds n + history()-1 items.  corr != NULL connects the optional second
// output (then the port-1 debug tags appear when two_ports is set); mag (optional) is filled
// from corr by the same magnitude kernel the block uses internally (d_corr_mag is private).
int ref_corr_est_work(void *h, int n, const float *in, uint64_t nitems_written, float *out0,
                      float *corr, float *mag, int two_ports, ao_tag *tags, int max_tags, int *ntags)
{
    gr::ais::corr_est_cc &b = **static_cast<gr::ais::corr_est_cc::sptr *>(h);
    *ntags = 0;
    if (n % b.output_multiple())
        return -1; // the scheduler never does this (set_output_multiple)
    std::vector<gr_complex> o0, o1;
    if (!out0) {
        o0.resize((size_t)n + 1);
        out0 = reinterpret_cast<float *>(o0.data());
    }
    bool second = two_ports || corr || mag;
    if (second && !corr) {
        o1.resize((size_t)n + 1);
        corr = reinterpret_cast<float *>(o1.data());
    }
    gr_vector_const_void_star iv = { in };
    gr_vector_void_star ov = { out0 };
    if (second)
        ov.push_back(corr);
    b.harness_set_counters(nitems_written, nitems_written);
    int r = b.work(n, iv, ov);
    if (mag)
        ao_mag_squared(corr, n, mag);
    // the block adds the tags of one detection as port0 x4 then port1 x3; the oracle lists them
    // in that order too, so merge the two per-port lists detection by detection
    std::vector<gr::tag_t> t0 = b.harness_take_added_tags(0), t1 = b.harness_take_added_tags(1);
    size_t i1 = 0;
    auto push = [&](const gr::tag_t &t, int port) {
        if (*ntags < max_tags) {
            tags[*ntags].offset = t.offset;
            tags[*ntags].key = tag_key(t.key);
            tags[*ntags].port = port;
            tags[*ntags].value = pmt::to_double(t.value);
        }
        (*ntags)++;
    };
    for (size_t i0 = 0; i0 < t0.size(); i0++) {
        push(t0[i0], 0);
        if (two_ports && (i0 % 4) == 3)
            for (int k = 0; k < 3 && i1 < t1.size(); k++)
                push(t1[i1++], 1);
    }
    return r;
}

// ---------------------------------------------------- msk_timing_recovery_cc
void *ref_msk_new(float sps, float gain, float limit, int osps, int *status)
{
    int rc = 0;
    RefMsk *m = 0;
    try {
        m = new RefMsk;
        m->blk = gr::ais::msk_timing_recovery_cc::make(sps, gain, limit, osps);
    } catch (const std::out_of_range &e) {
        // :82 "Gain must be positive" / :61 "osps must be 1 or 2" -- the oracle's -1 / -2
        rc = (std::string(e.what()).find("Gain") != std::string::npos) ? -1 : -2;
        delete m;
        m = 0;
    } catch (...) {
        rc = -9;
        delete m;
        m = 0;
    }
    if (status)
        *status = rc;
    return m;
}

void ref_msk_delete(void *h) { delete static_cast<RefMsk *>(h); }
float ref_msk_get_sps(void *h) { return static_cast<RefMsk *>(h)->blk->get_sps(); }
float ref_msk_get_gain(void *h) { return static_cast<RefMsk *>(h)->blk->get_gain(); }
float ref_msk_get_limit(void *h) { return static_cast<RefMsk *>(h)->blk->get_limit(); }
void ref_msk_set_sps(void *h, float v) { static_cast<RefMsk *>(h)->blk->set_sps(v); }
void ref_msk_set_limit(void *h, float v) { static_cast<RefMsk *>(h)->blk->set_limit(v); }
int ref_msk_set_gain(void *h, float v)
{
    try {
        static_cast<RefMsk *>(h)->blk->set_gain(v);
    } catch (const std::out_of_range &) {
        return -1;
    }
    return 0;
}
double ref_msk_relative_rate(void *h) { return static_cast<RefMsk *>(h)->blk->relative_rate(); }

int ref_msk_forecast(void *h, int noutput_items)
{
    gr_vector_int req(1, 0);
    static_cast<RefMsk *>(h)->blk->forecast(noutput_items, req);
    return req[0];
}

// One general_work() call.  tags: whatever travels on the input (all keys, any order); the
// block fetches its own "time_est" range.  The input is copied behind one item of history so
// that the block's in[-1] read after a negative centre (:148-153) sees what a GNU Radio buffer
// holds there: the last item consumed by the previous call (zero before the first).
int ref_msk_general_work(void *h, int noutput_items, int ninput_items, const float *in,
                         uint64_t nitems_read, const ao_tag *tags, int ntags, float *out,
                         float *out_err, float *out_mu, int *consumed)
{
    RefMsk *m = static_cast<RefMsk *>(h);
    m->buf.resize((size_t)ninput_items + 1);
    m->buf[0] = m->prev;
    if (ninput_items > 0)
        memcpy(&m->buf[1], in, sizeof(gr_complex) * (size_t)ninput_items);
    std::vector<gr::tag_t> tv;
    for (int t = 0; t < ntags; t++) {
        if (tags[t].port != 0)
            continue; // tags of corr_est's second output do not travel on output 0
        gr::tag_t g;
        g.offset = tags[t].offset;
        g.key = pmt::intern(key_name(tags[t].key));
        g.value = pmt::from_double(tags[t].value);
        tv.push_back(g);
    }
    m->blk->harness_set_input_tags(0, tv);
    m->blk->harness_set_counters(nitems_read, 0);
    m->blk->consume_each(0);
    gr_vector_int nin(1, ninput_items);
    gr_vector_const_void_star iv = { &m->buf[1] };
    std::vector<float> e, u;
    gr_vector_void_star ov = { out };
    // outputs are positional: connecting mu needs err connected too
    if (out_err || out_mu) {
        if (!out_err) {
            e.resize((size_t)noutput_items + 1);
            out_err = e.data();
        }
        ov.push_back(out_err);
        if (out_mu)
            ov.push_back(out_mu);
    }
    int k;
    try {
        k = m->blk->general_work(noutput_items, nin, iv, ov);
    } catch (const std::runtime_error &) {
        *consumed = 0;
        return -1; // the interpolator threw (mu out of range)
    }
    *consumed = m->blk->harness_consumed();
    if (*consumed > 0)
        m->prev = m->buf[(size_t)*consumed];
    return k;
}

// ------------------------------------------------------------------ freqest
void *ref_freqest_new(float sample_rate, int data_rate, int fftlen)
{
    try {
        return new gr::ais::freqest::sptr(gr::ais::freqest::make(sample_rate, data_rate, fftlen));
    } catch (...) {
        return 0;
    }
}
void ref_freqest_delete(void *h) { delete static_cast<gr::ais::freqest::sptr *>(h); }
int ref_freqest_work(void *h, const float *spec, int nvec, float *out)
{
    gr_vector_const_void_star iv = { spec };
    gr_vector_void_star ov = { out };
    return (*static_cast<gr::ais::freqest::sptr *>(h))->work(nvec, iv, ov);
}

// ------------------------------------------------------------------- invert
void ref_invert_work(const uint8_t *in, int n, uint8_t *out)
{
    static gr::ais::invert::sptr blk = gr::ais::invert::make(); // stateless
    gr_vector_const_void_star iv = { in };
    gr_vector_void_star ov = { out };
    blk->work(n, iv, ov);
}

// -------------------------------------------------------------- pdu_to_nmea
// posts a PDU (nil . blob) to the "to_nmea" port and returns what the block published on "out"
int ref_pdu_to_nmea(const char *designator, const uint8_t *data, int len, char *out, int cap)
{
    try {
        gr::ais::pdu_to_nmea::sptr blk = gr::ais::pdu_to_nmea::make(std::string(designator));
        blk->post(pmt::mp("to_nmea"), pmt::cons(pmt::PMT_NIL, pmt::make_blob(data, (size_t)len)));
        std::vector<pmt::pmt_t> &pub = blk->published("out");
        if (pub.size() != 1)
            return -1;
        pmt::pmt_t v = pmt::cdr(pub[0]);
        int n = (int)pmt::blob_length(v);
        if (n > cap)
            return -2;
        memcpy(out, pmt::blob_data(v), (size_t)n);
        return n;
    } catch (...) {
        return -3;
    }
}

// --------------------------------------- the block table for the oracle's chain
const ao_blocks *ref_blocks(void)
{
    static const ao_blocks b = { "reference",
                                 ref_corr_est_new,
                                 ref_corr_est_delete,
                                 ref_corr_est_output_multiple,
                                 ref_corr_est_set_symbols,
                                 ref_corr_est_work,
                                 ref_msk_new,
                                 ref_msk_delete,
                                 ref_msk_get_sps,
                                 ref_msk_general_work,
                                 ref_freqest_new,
                                 ref_freqest_delete,
                                 ref_freqest_work,
                                 ref_invert_work };
    return &b;
}

} // extern "C"
