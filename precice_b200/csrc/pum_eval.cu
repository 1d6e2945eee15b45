// pum_eval.cu — execution-mode "minimal-compute": the per-cluster evaluation matrices A_c(i, j) = phi(|out_i - in_j|)
// are computed once in computeMapping and stored; map() then streams them from HBM instead of re-evaluating the basis
// function (sm_100a).
//
// Reference: computeEvaluationOffline = true (PartitionOfUnityMapping.hpp:48-57, 145; XML attribute execution-mode,
// mapping/config/MappingConfiguration.cpp:328-331, 555-559): BatchedRBFSolver.hpp:196-239 allocates the evaluation
// matrices for all clusters and :324 fills them (kernel do_batched_assembly with the out-mesh as rows). The matrix of
// buildMatrixA (RadialBasisFctSolver.hpp:209-242) without polynomial columns (PUM only knows a separated polynomial).
//
// Layout: the map kernels assign one vertex of the LANE side to a thread and loop over the other (LOOP) side:
//   consistent  : lanes = out-vertices i, loop = in-vertices j   (out_i  = sum_j A(i,j) lambda_j)
//   conservative: lanes = in-vertices j,  loop = out-vertices i  (Au_j   = sum_i A(i,j) (w u)_i)
// so the entry of lane vertex l and loop vertex k is stored at evalOff[c] + k * ld + l, ld = number of lane vertices
// rounded up to even: consecutive threads read consecutive doubles for a fixed k (coalesced 8-byte loads).
#include "pum_device.cuh"

namespace rbfmap {
namespace {

constexpr int kEvalWarps = 4;

__global__ void evalSizeKernel(const int *__restrict__ nLane, const int *__restrict__ nLoop, int64_t nClusters, long long *__restrict__ size)
{
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c < nClusters)
    size[c] = static_cast<long long>((nLane[c] + 1) & ~1) * nLoop[c];
}

// one warp per cluster; the loop-side coordinates are staged in shared memory in chunks of kChunk vertices
template <int KIND, bool DIM3>
__global__ void __launch_bounds__(32 * kEvalWarps) pumEvalMatrixKernel(PumSolveArgs args, int64_t nClusters, bool lanesAreOut, const long long *__restrict__ evalOff,
                                                                       double *__restrict__ eval)
{
  constexpr int kChunk = 64;
  __shared__ __align__(16) double sx[kEvalWarps][kChunk], sy[kEvalWarps][kChunk], sz[kEvalWarps][kChunk];
  const int     w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c = blockIdx.x * static_cast<int64_t>(kEvalWarps) + w;
  if (c >= nClusters)
    return;
  const int       n = args.nIn[c], m = args.nOut[c];
  const int       nLane = lanesAreOut ? m : n, nLoop = lanesAreOut ? n : m;
  const int       ld = (nLane + 1) & ~1;
  const RbfParams rbf   = args.rbf;
  const double    invR2 = rbf.p1 * rbf.p1;
  const double   *inRec  = args.inRec + 3 * args.inOff[c];
  const double4  *outRec = args.outRec + args.outOff[c];
  double         *dst    = eval + evalOff[c];
  auto laneCoord = [&](int l, double &x, double &y, double &z) {
    if (lanesAreOut) {
      const double4 r = outRec[l];
      x = r.x, y = r.y, z = r.z;
    } else {
      x = inRec[3 * l], y = inRec[3 * l + 1], z = inRec[3 * l + 2];
    }
  };
  for (int k0 = 0; k0 < nLoop; k0 += kChunk) {
    const int kc = min(kChunk, nLoop - k0);
    __syncwarp();
    for (int k = lane; k < kc; k += 32) {
      double x, y, z;
      if (lanesAreOut) {
        x = inRec[3 * (k0 + k)], y = inRec[3 * (k0 + k) + 1], z = inRec[3 * (k0 + k) + 2];
      } else {
        const double4 r = outRec[k0 + k];
        x = r.x, y = r.y, z = r.z;
      }
      sx[w][k] = x, sy[w][k] = y, sz[w][k] = z;
    }
    __syncwarp();
    for (int l = lane; l < nLane; l += 32) {
      double x, y, z;
      laneCoord(l, x, y, z);
      double *col = dst + static_cast<long long>(k0) * ld + l;
      int     k   = 0;
      for (; k + 4 <= kc; k += 4) {
        double bx[4], by[4], bz[4], phi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          bx[u] = sx[w][k + u], by[u] = sy[w][k + u], bz[u] = sz[w][k + u];
        // the operand order of the map kernels: (out-vertex, in-vertex) for the consistent direction, (out, in) as well for
        // the conservative one (pairBasis(out..., in...) in pum_map_conservative.cu): the stored entry equals what they compute
        if (lanesAreOut) {
          pairBasisBatch<KIND, DIM3, 4>(x, y, z, bx, by, bz, rbf, invR2, phi);
        } else {
          double ax[4] = {x, x, x, x}, ay[4] = {y, y, y, y}, az[4] = {z, z, z, z};
          pairBasisBatch<KIND, DIM3, 4>(bx, by, bz, ax, ay, az, rbf, invR2, phi);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          col[static_cast<long long>(k + u) * ld] = phi[u];
      }
      for (; k < kc; ++k) {
        const double phi = lanesAreOut ? pairBasis<KIND, DIM3>(x, y, z, sx[w][k], sy[w][k], sz[w][k], rbf, invR2)
                                       : pairBasis<KIND, DIM3>(sx[w][k], sy[w][k], sz[w][k], x, y, z, rbf, invR2);
        col[static_cast<long long>(k) * ld] = phi;
      }
    }
  }
}

template <int KIND>
void launchEval(const PumSolveArgs &a, int64_t nClusters, bool lanesAreOut, const long long *evalOff, double *eval, cudaStream_t s)
{
  const int grid = static_cast<int>((nClusters + kEvalWarps - 1) / kEvalWarps);
  if (a.dim == 3)
    pumEvalMatrixKernel<KIND, true><<<grid, 32 * kEvalWarps, 0, s>>>(a, nClusters, lanesAreOut, evalOff, eval);
  else
    pumEvalMatrixKernel<KIND, false><<<grid, 32 * kEvalWarps, 0, s>>>(a, nClusters, lanesAreOut, evalOff, eval);
  RBF_LAUNCHED();
}

} // namespace

int64_t pumEvalOffsets(const PumMembership &m, int64_t nClusters, bool lanesAreOut, DevBuf<int64_t> &evalOff, cudaStream_t s)
{
  DevBuf<int64_t> size;
  size.alloc(nClusters);
  evalOff.alloc(nClusters + 1);
  evalSizeKernel<<<divUp(nClusters, 256), 256, 0, s>>>(lanesAreOut ? m.nOut.p : m.nIn.p, lanesAreOut ? m.nIn.p : m.nOut.p, nClusters,
                                                      reinterpret_cast<long long *>(size.p));
  RBF_LAUNCHED();
  exclusiveScan(size.p, evalOff.p, nClusters, s);
  int64_t total = 0;
  RBF_CUDA(cudaMemcpyAsync(&total, evalOff.p + nClusters, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  return total;
}

void pumBuildEvalMatrices(const PumSolveArgs &a, int64_t nClusters, bool lanesAreOut, const long long *evalOff, double *eval, cudaStream_t s)
{
  if (nClusters <= 0)
    return;
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, launchEval<KIND>(a, nClusters, lanesAreOut, evalOff, eval, s));
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
