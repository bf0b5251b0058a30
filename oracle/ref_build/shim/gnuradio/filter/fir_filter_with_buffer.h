// msk_timing_recovery_cc_impl.h:42 only declares a pointer to this kernel and never uses it.
#pragma once
namespace gr {
namespace filter {
namespace kernel {
class fir_filter_with_buffer_fff;
}
} // namespace filter
} // namespace gr
