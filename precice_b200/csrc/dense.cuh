// dense.cuh — dense FP64 building blocks of the global (single-system) RBF mappings:
//   * tiled assembly of the interpolation matrix (RadialBasisFctSolver.hpp:164-207, Kokkos K1 in SURVEY §2c)
//   * blocked Cholesky: register-tiled 64x64 diagonal factor + explicit inverse, panel and trailing updates as
//     FP64 tensor-core GEMMs (mma.sync m8n8k4 f64 = DMMA), replaces Eigen::LLT (RadialBasisFctSolver.hpp:352-353)
//   * blocked triangular solves with several right-hand sides (solveConsistent/solveConservative, :413-468)
//   * matrix-free products with the evaluation matrix (buildMatrixA, :209-242, never stored)
#pragma once

#include "basis.cuh"
#include "common.cuh"

namespace rbfmap {

struct AxisMask {
  double mx, my, mz; // 1.0 = active, 0.0 = dead (RadialBasisFctSolver.hpp:102-113)
  bool   allActive() const { return mx == 1.0 && my == 1.0 && mz == 1.0; }
};

constexpr int kDiagNb  = 64;  // diagonal block of the Cholesky / LU factorisations
constexpr int kPanelNb = 256; // outer panel = K of the trailing tensor-core update
inline int64_t padTo(int64_t n, int64_t b) { return (n + b - 1) / b * b; }

// C (np x np, column-major, leading dimension np) <- Phi(xyz, xyz) on the first n rows/cols, identity padding beyond.
// lowerOnly: only tiles on or below the diagonal are written.
void denseAssemble(const RbfParams &rbf, const AxisMask &mask, const double *dXyz, int64_t n, double *dC, int64_t np, bool lowerOnly, cudaStream_t s);

// In-place lower Cholesky factor of the np x np matrix (np multiple of kPanelNb); dInvDiag receives the explicit
// inverses of the 64x64 diagonal blocks (np/64 blocks of 64x64, column-major). Returns false if a pivot is <= 0.
bool denseCholesky(double *dC, int64_t np, double *dInvDiag, cudaStream_t s);

// Solves L L^T X = B in place for nrhs right-hand sides stored as separate vectors B[c*ldb + i] (ldb >= np).
void denseCholeskySolve(const double *dL, int64_t np, const double *dInvDiag, double *dB, int64_t ldb, int nrhs, cudaStream_t s);

// General (not positive definite) systems: in-place LU with partial pivoting of the np x np matrix (np multiple of 256),
// dPiv receives np row interchanges; returns false if numerically singular. Solve works on planar right-hand sides.
bool denseLu(double *dA, int64_t np, int *dPiv, cudaStream_t s);
void denseLuSolve(const double *dLU, int64_t np, const int *dPiv, double *dB, int64_t ldb, int nrhs, cudaStream_t s);

// rowOut[c*ldo + i] (=|+=) sum_j phi(|row_i - col_j|) colVals[c*ldv + j]  for i < nRows, j < nCols, c < nrhs
void denseEvaluate(const RbfParams &rbf, const AxisMask &mask, const double *dRowXyz, int64_t nRows, const double *dColXyz, int64_t nCols,
                   const double *dColVals, int64_t ldv, double *dRowOut, int64_t ldo, int nrhs, bool accumulate, cudaStream_t s);

// interleaved (vertex-major, values[i*dims + c]) <-> planar (c*ld + i) copies; planar rows beyond n are zero filled up to nFill
void interleavedToPlanar(const double *dIn, int64_t n, int dims, double *dOut, int64_t ld, int64_t nFill, cudaStream_t s);
void planarToInterleaved(const double *dIn, int64_t ld, int64_t n, int dims, double *dOut, cudaStream_t s);

} // namespace rbfmap
