// gsx_build.cu -- GPU index construction (suffix sorting on the device).  Placeholder until the builder lands.
#include "gsx_host.h"
namespace gsx {
bool build_strand_gpu(int, const uint8_t*, uint64_t, uint32_t, HostStrand&, std::string& err) {
    err = "gsx_index_build: GPU index construction is not implemented yet";
    return false;
}
}  // namespace gsx
