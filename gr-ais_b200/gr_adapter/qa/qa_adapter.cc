// qa_adapter.cc -- drives the four adapter blocks the way the GNU Radio scheduler would
// (stub runtime), on inputs read from a file, and dumps every output so that
// tests/test_gpu_adapter.py can compare them with the CPU oracle.
//
//   qa_adapter <in.bin> <out.bin>
// in.bin : int32 L, int32 n (items per work call, multiple of the block's output multiple),
//          int32 ncalls, L complex64 template, (ncalls*n + L) complex64 stream
// out.bin: see the writes below (all little-endian, fixed order)
#include <ais/corr_est_cc.h>
#include <ais/freqest.h>
#include <ais/invert.h>
#include <ais/pdu_to_nmea.h>
#include <ais/msk_timing_recovery_cc.h>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

static void put(FILE *f, const void *p, size_t bytes)
{
    if (bytes && fwrite(p, 1, bytes, f) != bytes)
        throw std::runtime_error("short write");
}

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
        return 2;
    }
    try {
        FILE *fi = fopen(argv[1], "rb");
        FILE *fo = fopen(argv[2], "wb");
        if (!fi || !fo)
            throw std::runtime_error("cannot open files");
        int32_t hdr[3];
        if (fread(hdr, sizeof(int32_t), 3, fi) != 3)
            throw std::runtime_error("bad header");
        const int L = hdr[0], n = hdr[1], ncalls = hdr[2];
        std::vector<gr_complex> tmpl((size_t)L), stream((size_t)ncalls * n + L);
        if (fread(tmpl.data(), sizeof(gr_complex), tmpl.size(), fi) != tmpl.size() ||
            fread(stream.data(), sizeof(gr_complex), stream.size(), fi) != stream.size())
            throw std::runtime_error("short read");
        fclose(fi);

        // corr_est_cc with both outputs connected -> msk_timing_recovery_cc with all three
        gr::ais::corr_est_cc::sptr ce = gr::ais::corr_est_cc::make(tmpl, 5.0f, 1, 0.9f);
        gr::ais::msk_timing_recovery_cc::sptr mk = gr::ais::msk_timing_recovery_cc::make(5.0f, 0.04f, 0.01f, 1);
        if ((int)ce->history() != L + 1 || n % ce->output_multiple())
            throw std::runtime_error("scheduler hints disagree with the test input");
        std::vector<gr_complex> out0((size_t)n), out1((size_t)n);
        std::vector<gr_complex> msk_in; // unconsumed corr_est output 0
        uint64_t msk_read = 0;
        std::vector<gr::tag_t> pending; // tags travelling downstream on output 0
        int32_t total_tags = 0, total_sym = 0;
        put(fo, &total_tags, 4); // patched at the end
        put(fo, &total_sym, 4);
        std::vector<gr_complex> sym;
        std::vector<float> err, mu;
        for (int call = 0; call < ncalls; call++) {
            gr_vector_const_void_star in = { stream.data() + (size_t)call * n };
            gr_vector_void_star out = { out0.data(), out1.data() };
            ce->harness_set_counters((uint64_t)call * n, (uint64_t)call * n);
            gr_vector_int nin = { n + L };
            if (ce->general_work(n, nin, in, out) != n)
                throw std::runtime_error("corr_est produced a short block");
            put(fo, out0.data(), sizeof(gr_complex) * n);
            put(fo, out1.data(), sizeof(gr_complex) * n);
            for (unsigned port = 0; port < 2; port++) {
                for (const gr::tag_t &t : ce->harness_take_added_tags(port)) {
                    int32_t p = (int32_t)port;
                    double v = pmt::to_double(t.value);
                    char key = pmt::symbol_to_string(t.key)[0] == 'p'   ? 1
                               : pmt::symbol_to_string(t.key)[0] == 't' ? 2
                               : pmt::symbol_to_string(t.key) == "corr_start" ? 0
                                                                               : 3;
                    int32_t k = key;
                    put(fo, &t.offset, 8);
                    put(fo, &k, 4);
                    put(fo, &p, 4);
                    put(fo, &v, 8);
                    total_tags++;
                    if (port == 0)
                        pending.push_back(t);
                }
            }
            // hand everything produced so far to the timing-recovery block
            msk_in.insert(msk_in.end(), out0.begin(), out0.end());
            const int avail = (int)msk_in.size();
            const int want = avail; // more than enough room
            std::vector<gr_complex> o((size_t)want + 1);
            std::vector<float> e((size_t)want + 1), m((size_t)want + 1);
            gr_vector_const_void_star min = { msk_in.data() };
            gr_vector_void_star mout = { o.data(), e.data(), m.data() };
            gr_vector_int mnin = { avail };
            mk->harness_set_counters(msk_read, 0);
            mk->harness_set_input_tags(0, pending);
            const int k = mk->general_work(want, mnin, min, mout);
            const int c = mk->harness_consumed();
            sym.insert(sym.end(), o.begin(), o.begin() + k);
            err.insert(err.end(), e.begin(), e.begin() + k);
            mu.insert(mu.end(), m.begin(), m.begin() + k);
            msk_in.erase(msk_in.begin(), msk_in.begin() + c);
            msk_read += (uint64_t)c;
        }
        total_sym = (int32_t)sym.size();
        put(fo, sym.data(), sizeof(gr_complex) * sym.size());
        put(fo, err.data(), sizeof(float) * err.size());
        put(fo, mu.data(), sizeof(float) * mu.size());

        // freqest on the first 1024-item vectors of the stream, invert on some bytes
        const int nvec = (int)(stream.size() / 1024);
        gr::ais::freqest::sptr fe = gr::ais::freqest::make(48000.0f, 9600, 1024);
        std::vector<float> hz((size_t)nvec);
        gr_vector_const_void_star fin = { stream.data() };
        gr_vector_void_star fout = { hz.data() };
        fe->work(nvec, fin, fout);
        int32_t nv = nvec;
        put(fo, &nv, 4);
        put(fo, hz.data(), sizeof(float) * hz.size());
        std::vector<char> bytes(1000), inv(1000);
        for (size_t i = 0; i < bytes.size(); i++)
            bytes[i] = (char)(i * 7 + 3);
        gr::ais::invert::sptr iv = gr::ais::invert::make();
        gr_vector_const_void_star iin = { bytes.data() };
        gr_vector_void_star iout = { inv.data() };
        iv->work((int)bytes.size(), iin, iout);
        put(fo, inv.data(), inv.size());

        // argument errors surface as the reference's exceptions
        int32_t threw = 0;
        try {
            gr::ais::msk_timing_recovery_cc::make(5.0f, 0.0f, 0.01f, 1);
        } catch (const std::out_of_range &) {
            threw |= 1;
        }
        try {
            gr::ais::msk_timing_recovery_cc::make(5.0f, 0.04f, 0.01f, 3);
        } catch (const std::out_of_range &) {
            threw |= 2;
        }
        put(fo, &threw, 4);

        // pdu_to_nmea: a 21-byte PDU (bytes 7*i+1) and a 53-byte one posted to "to_nmea";
        // the sentences come back as u8vector PDUs on "out"
        gr::ais::pdu_to_nmea::sptr nm = gr::ais::pdu_to_nmea::make("B");
        for (int len : { 21, 53 }) {
            std::vector<uint8_t> pdu((size_t)len);
            for (int i = 0; i < len; i++)
                pdu[(size_t)i] = (uint8_t)(7 * i + 1);
            nm->post(pmt::mp("to_nmea"), pmt::cons(pmt::PMT_NIL, pmt::make_blob(pdu.data(), pdu.size())));
        }
        for (auto &m : nm->published("out")) {
            int32_t sl = (int32_t)pmt::blob_length(pmt::cdr(m));
            put(fo, &sl, 4);
            put(fo, pmt::blob_data(pmt::cdr(m)), (size_t)sl);
        }
        fseek(fo, 0, SEEK_SET);
        put(fo, &total_tags, 4);
        put(fo, &total_sym, 4);
        fclose(fo);
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "qa_adapter: %s\n", e.what());
        return 1;
    }
}
