#pragma once
#include <gnuradio/block.h>
