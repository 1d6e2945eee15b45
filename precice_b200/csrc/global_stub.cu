// placeholder until global_direct.cu / global_iterative.cu land
#include "mapping.cuh"
namespace rbfmap {
std::unique_ptr<DeviceMapping> makeGlobalDirectMapping(const rbfmap_config &) { throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-direct: not built yet"); }
std::unique_ptr<DeviceMapping> makeGlobalIterativeMapping(const rbfmap_config &) { throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-iterative: not built yet"); }
} // namespace rbfmap
