// stub_runtime.h -- the handful of GNU Radio 3.8 runtime types the four AIS blocks touch,
// reduced to what a block needs to compile and to be driven by a test harness.
//
// GNU Radio is not installed in the build environment, so the adapter
// in ../lib is compile-checked and exercised against this stub; with -DB200AIS_HAVE_GNURADIO
// the same adapter sources include the real <gnuradio/...> headers instead.  Only semantics
// the blocks rely on are modelled: item counters, history, stream tags, the virtual
// work()/general_work()/forecast() entry points with GNU Radio's exact signatures.
#pragma once

#include <algorithm>
#include <complex>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}

namespace pmt {
struct pmt_base;
typedef std::shared_ptr<pmt_base> pmt_t;
struct pmt_base {
    bool is_symbol = false;
    std::string sym;
    double dbl = 0.0;
    std::vector<uint8_t> blob; // blob / u8vector payload
    pmt_t car, cdr;            // pair
};
inline pmt_t intern(const std::string &s)
{
    static std::map<std::string, pmt_t> table;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto it = table.find(s);
    if (it != table.end())
        return it->second;
    pmt_t p = std::make_shared<pmt_base>();
    p->is_symbol = true;
    p->sym = s;
    table[s] = p;
    return p;
}
inline pmt_t from_double(double v)
{
    pmt_t p = std::make_shared<pmt_base>();
    p->dbl = v;
    return p;
}
inline double to_double(const pmt_t &p) { return p->dbl; }
inline bool eqv(const pmt_t &a, const pmt_t &b) { return a == b; }
inline std::string symbol_to_string(const pmt_t &p) { return p->sym; }
// the PDU subset gr::ais::pdu_to_nmea uses (lib/pdu_to_nmea_impl.cc:64-65,139-140)
static const pmt_t PMT_NIL = std::make_shared<pmt_base>();
inline pmt_t mp(const std::string &s) { return intern(s); }
inline pmt_t cons(const pmt_t &a, const pmt_t &b)
{
    pmt_t p = std::make_shared<pmt_base>();
    p->car = a;
    p->cdr = b;
    return p;
}
inline pmt_t car(const pmt_t &p) { return p->car; }
inline pmt_t cdr(const pmt_t &p) { return p->cdr; }
inline pmt_t make_blob(const void *data, size_t len)
{
    pmt_t p = std::make_shared<pmt_base>();
    p->blob.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + len);
    return p;
}
inline pmt_t init_u8vector(size_t len, const uint8_t *data) { return make_blob(data, len); }
inline const void *blob_data(const pmt_t &p) { return p->blob.data(); }
inline size_t blob_length(const pmt_t &p) { return p->blob.size(); }
} // namespace pmt

namespace gr {

struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value, srcid;
};

namespace thread {
typedef std::mutex mutex;
typedef std::lock_guard<std::mutex> scoped_lock;
} // namespace thread

class io_signature
{
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    {
        return sptr(new io_signature(min_streams, max_streams, { sizeof_stream_item }));
    }
    static sptr make3(int min_streams, int max_streams, int s1, int s2, int s3)
    {
        return sptr(new io_signature(min_streams, max_streams, { s1, s2, s3 }));
    }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int i) const
    {
        return d_sizes[std::min<size_t>((size_t)i, d_sizes.size() - 1)];
    }

private:
    io_signature(int mn, int mx, std::vector<int> sizes) : d_min(mn), d_max(mx), d_sizes(sizes) {}
    int d_min, d_max;
    std::vector<int> d_sizes;
};

class basic_block
{
public:
    virtual ~basic_block() {}
    std::string name() const { return d_name; }
    std::string alias() const { return d_name + "0"; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }

protected:
    basic_block() {}
    basic_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_in(in), d_out(out)
    {
    }
    std::string d_name;
    io_signature::sptr d_in, d_out;

    // message ports: handlers run synchronously in post(); published messages are kept for
    // the harness to read
public:
    typedef std::function<void(pmt::pmt_t)> msg_handler_t;
    void message_port_register_in(pmt::pmt_t port) { d_handlers[pmt::symbol_to_string(port)]; }
    void message_port_register_out(pmt::pmt_t port) { d_published[pmt::symbol_to_string(port)]; }
    void set_msg_handler(pmt::pmt_t port, msg_handler_t h) { d_handlers[pmt::symbol_to_string(port)] = h; }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg)
    {
        d_published[pmt::symbol_to_string(port)].push_back(msg);
    }
    void post(pmt::pmt_t port, pmt::pmt_t msg) { d_handlers.at(pmt::symbol_to_string(port))(msg); }
    std::vector<pmt::pmt_t> &published(const std::string &port) { return d_published[port]; }

private:
    std::map<std::string, msg_handler_t> d_handlers;
    std::map<std::string, std::vector<pmt::pmt_t>> d_published;
};

// gr::block: general_work() with forecast(); the harness sets the item counters and the
// tags visible on the inputs, and reads back what the block added and consumed.
class block : public basic_block
{
public:
    virtual void forecast(int noutput_items, gr_vector_int &ninput_items_required)
    {
        for (auto &v : ninput_items_required)
            v = noutput_items + (int)history() - 1;
    }
    // not pure, as in GNU Radio (message-only blocks such as pdu_to_nmea do not override it)
    virtual int general_work(int, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &)
    {
        throw std::runtime_error("block::general_work() not implemented");
    }

    unsigned history() const { return d_history; }
    void set_history(unsigned h) { d_history = h; }
    int output_multiple() const { return d_output_multiple; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    double relative_rate() const { return d_relative_rate; }
    void set_relative_rate(double r) { d_relative_rate = r; }
    void enable_update_rate(bool en) { d_update_rate = en; }
    void declare_sample_delay(int which, unsigned delay) { d_sample_delay[which] = delay; }
    unsigned sample_delay(int which) const
    {
        auto it = d_sample_delay.find(which);
        return it == d_sample_delay.end() ? 0u : it->second;
    }
    int max_noutput_items() const { return d_max_noutput; }
    void set_max_noutput_items(int m) { d_max_noutput = m; }
    void consume_each(int n) { d_consumed = n; }
    uint64_t nitems_read(unsigned) const { return d_nitems_read; }
    uint64_t nitems_written(unsigned) const { return d_nitems_written; }
    void add_item_tag(unsigned port, uint64_t offset, const pmt::pmt_t &key, const pmt::pmt_t &value,
                      const pmt::pmt_t &srcid = pmt::pmt_t())
    {
        tag_t t;
        t.offset = offset;
        t.key = key;
        t.value = value;
        t.srcid = srcid;
        d_added_tags[port].push_back(t);
    }
    void get_tags_in_range(std::vector<tag_t> &v, unsigned port, uint64_t start, uint64_t end,
                           const pmt::pmt_t &key)
    {
        v.clear();
        for (const auto &t : d_input_tags[port])
            if (t.offset >= start && t.offset < end && pmt::eqv(t.key, key))
                v.push_back(t);
        std::stable_sort(v.begin(), v.end(),
                         [](const tag_t &a, const tag_t &b) { return a.offset < b.offset; });
    }

    // ---- harness side (what the scheduler would own) ----
    void harness_set_counters(uint64_t nread, uint64_t nwritten)
    {
        d_nitems_read = nread;
        d_nitems_written = nwritten;
    }
    void harness_set_input_tags(unsigned port, const std::vector<tag_t> &tags) { d_input_tags[port] = tags; }
    std::vector<tag_t> harness_take_added_tags(unsigned port)
    {
        std::vector<tag_t> r;
        r.swap(d_added_tags[port]);
        return r;
    }
    int harness_consumed() const { return d_consumed; }

protected:
    block() {}
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : basic_block(name, in, out)
    {
    }
    gr::thread::mutex d_setlock; // gr::block's own member (corr_est_cc_impl.cc:135,169 lock it)

private:
    unsigned d_history = 1;
    int d_output_multiple = 1;
    double d_relative_rate = 1.0;
    bool d_update_rate = false;
    std::map<int, unsigned> d_sample_delay;
    int d_max_noutput = 0;
    int d_consumed = 0;
    uint64_t d_nitems_read = 0, d_nitems_written = 0;
    std::map<unsigned, std::vector<tag_t>> d_added_tags, d_input_tags;
};

class sync_block : public block
{
public:
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) = 0;
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override
    {
        int r = work(noutput_items, input_items, output_items);
        if (r > 0)
            consume_each(r);
        return r;
    }

protected:
    sync_block() {}
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : block(name, in, out)
    {
    }
};

} // namespace gr

namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T *p) { return std::shared_ptr<T>(p); }
} // namespace gnuradio

#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT
