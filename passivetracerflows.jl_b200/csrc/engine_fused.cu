// Fused engine placeholder: replaced by the hand-written shared-memory FFT pipeline (engine_fused2d.cu).
#include "ptf_internal.h"
namespace ptf {
bool fused_engine_supports(const Context&, std::string* why) {
  if (why) *why = "fused engine not built";
  return false;
}
std::unique_ptr<Engine> make_fused_engine(Context&) { throw Error(PTF_EUNSUPPORTED, "fused engine not built"); }
}  // namespace ptf
