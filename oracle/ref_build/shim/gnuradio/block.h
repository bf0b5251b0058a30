// Shim of <gnuradio/block.h> for compiling the gr-ais reference sources UNMODIFIED without
// GNU Radio (oracle/ref_build/Makefile).  The runtime types are the in-tree stub the block
// adapters are compiled against; on top of it, the few Boost names GNU Radio 3.8's own headers
// bring into scope and that lib/pdu_to_nmea_impl.cc:52-54,106-113 uses unqualified.
// TEST INFRASTRUCTURE ONLY (oracle/).
#pragma once
#include <gnuradio/stub_runtime.h>

#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace boost {
using std::bind;
// boost/exception/to_string.hpp: operator<< into a stringstream
template <class T> inline std::string to_string(const T &v)
{
    std::ostringstream o;
    o << v;
    return o.str();
}
} // namespace boost
using std::placeholders::_1; // boost/bind.hpp puts the placeholders in the global namespace
