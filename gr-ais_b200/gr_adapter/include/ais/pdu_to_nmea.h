// gr::ais::pdu_to_nmea, B200 build (reference include/ais/pdu_to_nmea.h:37-54): a message
// block -- PDUs (pmt pair with a blob in the cdr) arrive on "print" / "to_nmea"; "print" writes
// the AIVDM sentence to stdout, "to_nmea" publishes it as a u8vector PDU on "out".
#ifndef B200AIS_GR_AIS_PDU_TO_NMEA_H
#define B200AIS_GR_AIS_PDU_TO_NMEA_H

#include <ais/api.h>
#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/block.h>
#endif
#include <string>

namespace gr {
namespace ais {

class AIS_API pdu_to_nmea : virtual public gr::block
{
public:
    typedef boost::shared_ptr<pdu_to_nmea> sptr;
    virtual void to_nmea(pmt::pmt_t) = 0;
    virtual void print(pmt::pmt_t) = 0;
    static sptr make(std::string designator);
};

} // namespace ais
} // namespace gr

#endif
