// placeholder until global_iterative.cu lands
#include "mapping.cuh"
namespace rbfmap {
std::unique_ptr<DeviceMapping> makeGlobalIterativeMapping(const rbfmap_config &) { throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-iterative: not built yet"); }
} // namespace rbfmap
