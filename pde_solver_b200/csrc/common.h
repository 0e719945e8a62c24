// Internal definitions shared by the CUDA translation units of libpde_b200.so.
//
// HBM layout of a nodal field (one displacement/temperature component):
//   node (ix,iy,iz) lives at  base + ix + PX*(iy + PY*iz)
//   PX >= nnx+1 (multiple of 4 doubles: 32-byte aligned rows, TMA-legal pitch), PY = nny+1
//   (PY = 1 when the y axis is absent).  PDE_NG zero "ghost" planes sit before plane 0 and
//   after the last plane; every pad column / pad row / ghost plane stays ZERO in all vectors,
//   so a stencil read that leaves the domain (including the x/y wrap into the previous row /
//   plane) reads 0 and no bounds test is needed.  With slab partitioning the two ghost planes
//   are the halo planes received from the z-neighbours.
//   2D problems use internal axes (x, z): a "plane" is one row, so slabs/halos work unchanged.
//   Vector problems are SoA: component c at base + c*comp_stride.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/pde_b200.h"

// ghost planes kept before plane 0 and after the last plane of every field component (zero at the domain
// ends, halo planes between slabs).  Depth 2 lets the temporally blocked smoother run two sweeps per pass.
#define PDE_NG 2
#define PDE_NCLASS 27
#define PDE_NOFF 15

struct Grid {
  int dim;          // user dimension 1..3
  int nn[3];        // nodes per internal axis (1 if absent)
  int nc[3];        // cells per internal axis (0 if absent)
  int PX, PY;       // pitches
  int nzl;          // local planes owned by this rank
  int z0;           // global index of first local plane
  int nzg;          // global planes
  long long plane;  // PX*PY
  long long total;  // plane*nzl : flat length of one component (without ghost planes)
  long long comp_stride;  // plane*(nzl+2*PDE_NG)
  double h[3];
  int nk;                 // active stencil offsets
  int kidx[PDE_NOFF];     // active offset ids
  long long koff[PDE_NOFF];  // flat offsets for active ids
  int kdx[PDE_NOFF], kdy[PDE_NOFF], kdz[PDE_NOFF];
};

struct BcDev {
  int on[6];     // internal faces: x0,x1,y0,y1,z0,z1
  double val[6];
  int side_excl;
};

// Offsets of the Kuhn/Freudenthal 15-point stencil (internal axes).
static const int kOffD[PDE_NOFF][3] = {
    {0, 0, 0},  {1, 0, 0},  {-1, 0, 0},  {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1}, {1, 1, 0},
    {-1, -1, 0}, {1, 0, 1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, -1}, {1, 1, 1}, {-1, -1, -1}};

// thread-local error string
void pde_set_error(const std::string& s);
#define PDE_FAIL(msg)                                                          \
  do {                                                                         \
    pde_set_error(std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    return 1;                                                                  \
  } while (0)
#define CUDA_OK(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      pde_set_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + #call + " (" + \
                    __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
      return 1;                                                                \
    }                                                                          \
  } while (0)
#define PDE_OK(call)           \
  do {                         \
    int r__ = (call);          \
    if (r__) return r__;       \
  } while (0)

// ---- host-side stencil tables (tables.cpp) ------------------------------------------------
struct OpTable {
  int ncomp = 1;
  std::vector<double> coef;   // [27][15][ncomp*ncomp]
  std::vector<double> load;   // [27] : m_i = integral of the hat function
  double gershgorin = 0;      // max_i sum_j |a_ij| / a_ii over classes (upper bound of lmax(D^-1 A))
};
// axes present: dim 1 -> {x}; dim 2 -> {x,z}; dim 3 -> {x,y,z}
void internal_axes(int dim, int ax[3], int* nax);
int build_scalar_table(int dim, const double h_int[3], const int nc_int[3], double alpha, double beta,
                       OpTable* out);
int build_elasticity_table(int dim, const double h_int[3], const int nc_int[3], double lam, double mu,
                           OpTable* out);
// simplex geometry of the reference cell split (for the von-Mises kernel)
struct SimplexGeom {
  int nsimp;          // simplices per grid cell
  int nv;             // vertices per simplex
  int corner[6][4];   // corner bitmasks over internal axes (bit0=x, bit1=y, bit2=z)
  double G[6][4][3];  // basis gradients, internal axes
  double vol;         // simplex volume (all equal)
};
void build_simplex_geom(int dim, const double h_int[3], SimplexGeom* out);

int make_grid(int dim, const int32_t n_user[3], const double L_user[3], int rank, int world, Grid* g);
void user_bc_to_dev(int dim, const pde_bc* bc, BcDev* out);
