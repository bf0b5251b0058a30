// Symbol visibility for the B200 build of the gr-ais demod blocks (mirrors the role of the
// reference's include/ais/api.h:27-31).
#ifndef B200AIS_GR_AIS_API_H
#define B200AIS_GR_AIS_API_H

#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/attributes.h>
#else
#include <gnuradio/stub_runtime.h>
#endif

#ifdef gnuradio_ais_EXPORTS
#define AIS_API __GR_ATTR_EXPORT
#else
#define AIS_API __GR_ATTR_IMPORT
#endif

#endif
