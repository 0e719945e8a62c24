// Host-side construction of the matrix-free operator tables.
//
// On a uniform Interval/Rectangle("right")/Box mesh (reference: fenics_mcp_server.py:229,369,533)
// every grid cell carries the same simplices, so row i of any P1 operator is
//     sum over the <= 2^d grid cells around node i that exist,
//     sum over the simplices of that cell that contain i,   Ke[local(i)][local(j)] * x_j .
// Which cells exist depends only on whether the node sits on the low face / inside / on the
// high face of each axis: 27 "classes".  coef[class][offset][ci*nc+cj] holds the summed element
// entries, offset being one of the 15 Kuhn-stencil neighbours.  Class 13 (interior) is the
// constant stencil of SURVEY appendix A.2; the other classes are the partial patches on
// traction-free / natural boundaries.  Element matrices are the exact P1 integrals the
// reference gets from FFC (mass: degree-2 exact, stiffness/elasticity: constant integrand).
#include <cmath>
#include <cstring>

#include "common.h"

static thread_local std::string g_err;
void pde_set_error(const std::string& s) { g_err = s; }
extern "C" const char* pde_last_error(void) { return g_err.c_str(); }

void internal_axes(int dim, int ax[3], int* nax) {
  if (dim == 1) { ax[0] = 0; *nax = 1; }
  else if (dim == 2) { ax[0] = 0; ax[1] = 2; *nax = 2; }
  else { ax[0] = 0; ax[1] = 1; ax[2] = 2; *nax = 3; }
}

static int off_index(int dx, int dy, int dz) {
  for (int k = 0; k < PDE_NOFF; ++k)
    if (kOffD[k][0] == dx && kOffD[k][1] == dy && kOffD[k][2] == dz) return k;
  return -1;
}

// DOLFIN cell splits, corners as user-axis bitmasks (bit0 = x, bit1 = y, bit2 = z):
//   interval (v0,v1); rectangle "right" (v0,v1,v3),(v0,v2,v3); box: six tets sharing v0-v7.
static const int kSimp1[1][2] = {{0, 1}};
static const int kSimp2[2][3] = {{0, 1, 3}, {0, 2, 3}};
static const int kSimp3[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4},
                                 {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};

void build_simplex_geom(int dim, const double h[3], SimplexGeom* sg) {
  int ax[3], nax;
  internal_axes(dim, ax, &nax);
  sg->nsimp = dim == 1 ? 1 : (dim == 2 ? 2 : 6);
  sg->nv = dim + 1;
  std::memset(sg->G, 0, sizeof(sg->G));
  for (int t = 0; t < sg->nsimp; ++t) {
    for (int a = 0; a <= dim; ++a) {
      int ub = dim == 1 ? kSimp1[t][a] : (dim == 2 ? kSimp2[t][a] : kSimp3[t][a]);
      int ib = 0;
      for (int q = 0; q < nax; ++q)
        if (ub & (1 << q)) ib |= 1 << ax[q];
      sg->corner[t][a] = ib;
    }
    // edge matrix E[r][q] = (p_{r+1} - p_0)[axis q];  grad phi_a (a>=1) = column a-1 of E^-1
    double E[3][3] = {{0}}, Inv[3][3] = {{0}};
    for (int r = 0; r < dim; ++r)
      for (int q = 0; q < dim; ++q) {
        int b1 = (sg->corner[t][r + 1] >> ax[q]) & 1, b0 = (sg->corner[t][0] >> ax[q]) & 1;
        E[r][q] = (b1 - b0) * h[ax[q]];
      }
    // Gauss-Jordan inverse with partial pivoting (d <= 3)
    double A[3][6];
    for (int r = 0; r < dim; ++r)
      for (int q = 0; q < dim; ++q) { A[r][q] = E[r][q]; A[r][dim + q] = (r == q); }
    for (int c = 0; c < dim; ++c) {
      int piv = c;
      for (int r = c + 1; r < dim; ++r)
        if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
      for (int q = 0; q < 2 * dim; ++q) std::swap(A[c][q], A[piv][q]);
      double d = A[c][c];
      for (int q = 0; q < 2 * dim; ++q) A[c][q] /= d;
      for (int r = 0; r < dim; ++r)
        if (r != c) {
          double f = A[r][c];
          for (int q = 0; q < 2 * dim; ++q) A[r][q] -= f * A[c][q];
        }
    }
    for (int r = 0; r < dim; ++r)
      for (int q = 0; q < dim; ++q) Inv[r][q] = A[r][dim + q];
    for (int a = 1; a <= dim; ++a)
      for (int q = 0; q < dim; ++q) sg->G[t][a][ax[q]] = Inv[q][a - 1];
    for (int q = 0; q < dim; ++q) {
      double s = 0;
      for (int a = 1; a <= dim; ++a) s += sg->G[t][a][ax[q]];
      sg->G[t][0][ax[q]] = -s;
    }
  }
  double v = 1.0;
  for (int q = 0; q < nax; ++q) v *= h[ax[q]];
  sg->vol = v / (dim == 1 ? 1.0 : (dim == 2 ? 2.0 : 6.0));
}

// generic accumulation: elem(t, a, b, out[ncomp*ncomp])
template <class ElemFn>
static int build_table(int dim, const double h[3], int ncomp, ElemFn elem, OpTable* out) {
  int ax[3], nax;
  internal_axes(dim, ax, &nax);
  SimplexGeom sg;
  build_simplex_geom(dim, h, &sg);
  const int nn = ncomp * ncomp;
  out->ncomp = ncomp;
  out->coef.assign((size_t)PDE_NCLASS * PDE_NOFF * nn, 0.0);
  out->load.assign(PDE_NCLASS, 0.0);
  int present[3] = {0, 0, 0};
  for (int q = 0; q < nax; ++q) present[ax[q]] = 1;
  std::vector<double> e(nn);
  for (int cls = 0; cls < PDE_NCLASS; ++cls) {
    int c[3] = {cls % 3, (cls / 3) % 3, cls / 9};
    bool skip = false;
    for (int k = 0; k < 3; ++k)
      if (!present[k] && c[k] != 1) skip = true;  // absent axes only have the "mid" class
    if (skip) continue;
    for (int o = 0; o < 8; ++o) {
      bool valid = true;
      for (int k = 0; k < 3; ++k) {
        int bit = (o >> k) & 1;
        if (!present[k]) { if (bit) valid = false; continue; }
        if (bit && c[k] == 0) valid = false;   // cell on the minus side needs a lower neighbour
        if (!bit && c[k] == 2) valid = false;  // cell on the plus side needs an upper neighbour
      }
      if (!valid) continue;
      for (int t = 0; t < sg.nsimp; ++t)
        for (int a = 0; a < sg.nv; ++a) {
          if (sg.corner[t][a] != o) continue;
          out->load[cls] += sg.vol / (dim + 1);
          for (int b = 0; b < sg.nv; ++b) {
            int cb = sg.corner[t][b];
            int d[3];
            for (int k = 0; k < 3; ++k) d[k] = ((cb >> k) & 1) - ((o >> k) & 1);
            int kk = off_index(d[0], d[1], d[2]);
            if (kk < 0) { pde_set_error("stencil offset outside the Kuhn set"); return 1; }
            elem(sg, t, a, b, e.data());
            double* dst = &out->coef[((size_t)cls * PDE_NOFF + kk) * nn];
            for (int q = 0; q < nn; ++q) dst[q] += e[q];
          }
        }
    }
  }
  // Gershgorin bound of lmax(D^-1 A)
  double g = 0;
  for (int cls = 0; cls < PDE_NCLASS; ++cls)
    for (int i = 0; i < ncomp; ++i) {
      double diag = out->coef[((size_t)cls * PDE_NOFF + 0) * nn + i * ncomp + i];
      if (diag <= 0) continue;
      double s = 0;
      for (int k = 0; k < PDE_NOFF; ++k)
        for (int j = 0; j < ncomp; ++j)
          s += std::fabs(out->coef[((size_t)cls * PDE_NOFF + k) * nn + i * ncomp + j]);
      if (s / diag > g) g = s / diag;
    }
  out->gershgorin = g;
  return 0;
}

int build_scalar_table(int dim, const double h[3], const int nc[3], double alpha, double beta,
                       OpTable* out) {
  (void)nc;
  auto elem = [=](const SimplexGeom& sg, int t, int a, int b, double* e) {
    double gg = 0;
    for (int k = 0; k < 3; ++k) gg += sg.G[t][a][k] * sg.G[t][b][k];
    double m = sg.vol / ((dim + 1) * (dim + 2)) * (a == b ? 2.0 : 1.0);
    e[0] = alpha * m + beta * sg.vol * gg;
  };
  return build_table(dim, h, 1, elem, out);
}

int build_elasticity_table(int dim, const double h[3], const int nc[3], double lam, double mu,
                           OpTable* out) {
  (void)nc;
  int ax[3], nax;
  internal_axes(dim, ax, &nax);
  const int ncomp = dim;
  // component ci <-> internal axis ax[ci]
  auto elem = [=](const SimplexGeom& sg, int t, int a, int b, double* e) {
    double gg = 0;
    for (int k = 0; k < 3; ++k) gg += sg.G[t][a][k] * sg.G[t][b][k];
    for (int i = 0; i < ncomp; ++i)
      for (int j = 0; j < ncomp; ++j) {
        double gai = sg.G[t][a][ax[i]], gbj = sg.G[t][b][ax[j]];
        double gaj = sg.G[t][a][ax[j]], gbi = sg.G[t][b][ax[i]];
        e[i * ncomp + j] = sg.vol * (lam * gai * gbj + mu * gaj * gbi + (i == j ? mu * gg : 0.0));
      }
  };
  return build_table(dim, h, ncomp, elem, out);
}

int make_grid(int dim, const int32_t n[3], const double L[3], int rank, int world, Grid* g) {
  if (dim < 1 || dim > 3) PDE_FAIL("dim must be 1, 2 or 3");
  for (int k = 0; k < dim; ++k) {
    if (n[k] < 1) PDE_FAIL("cell counts must be >= 1");
    if (!(L[k] > 0)) PDE_FAIL("domain lengths must be > 0");
  }
  std::memset(g, 0, sizeof(*g));
  g->dim = dim;
  int ax[3], nax;
  internal_axes(dim, ax, &nax);
  for (int k = 0; k < 3; ++k) { g->nc[k] = 0; g->nn[k] = 1; g->h[k] = 1.0; }
  for (int q = 0; q < nax; ++q) {
    g->nc[ax[q]] = n[q];
    g->nn[ax[q]] = n[q] + 1;
    g->h[ax[q]] = L[q] / (double)n[q];
  }
  g->PX = ((g->nn[0] + 1 + 3) / 4) * 4;
  g->PY = g->nc[1] > 0 ? g->nn[1] + 1 : 1;
  g->plane = (long long)g->PX * g->PY;
  g->nzg = g->nn[2];
  if (world > 1) {
    if (g->nc[2] < world) PDE_FAIL("slab partition needs at least 1 cell layer per GPU along the slowest axis");
    int base = g->nc[2] / world;
    g->z0 = rank * base;
    g->nzl = (rank == world - 1) ? (g->nzg - g->z0) : base;
  } else {
    g->z0 = 0;
    g->nzl = g->nzg;
  }
  g->total = g->plane * g->nzl;
  g->comp_stride = g->plane * (g->nzl + 2 * PDE_NG);
  const int act1[] = {0, 1, 2}, act2[] = {0, 1, 2, 5, 6, 9, 10};
  if (dim == 1) { g->nk = 3; for (int k = 0; k < 3; ++k) g->kidx[k] = act1[k]; }
  else if (dim == 2) { g->nk = 7; for (int k = 0; k < 7; ++k) g->kidx[k] = act2[k]; }
  else { g->nk = 15; for (int k = 0; k < 15; ++k) g->kidx[k] = k; }
  for (int k = 0; k < g->nk; ++k) {
    const int* d = kOffD[g->kidx[k]];
    g->kdx[k] = d[0]; g->kdy[k] = d[1]; g->kdz[k] = d[2];
    g->koff[k] = d[0] + (long long)g->PX * d[1] + g->plane * d[2];
  }
  return 0;
}

void user_bc_to_dev(int dim, const pde_bc* bc, BcDev* out) {
  std::memset(out, 0, sizeof(*out));
  if (!bc) return;
  // user faces: x0,x1,(y0,y1),(z0,z1) of the user's axes; dim 2 maps user y -> internal z
  for (int f = 0; f < 2 * dim; ++f) {
    int uax = f / 2, side = f % 2;
    int iax = (dim == 2 && uax == 1) ? 2 : uax;
    out->on[2 * iax + side] = bc->face_on[f];
    out->val[2 * iax + side] = bc->face_val[f];
  }
  out->side_excl = bc->side_excludes_xends;
}

extern "C" int pde_version(void) { return 100; }

extern "C" void pde_solver_opts_default(pde_solver_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->rtol = 1e-10;
  o->max_iters = 100000;
  o->precond = PDE_PRECOND_AUTO;
  o->cheby_degree = 2;
  o->check_every = 10;
  o->cheby_ratio = 8.0;
}

extern "C" int pde_mesh_counts(int dim, const int32_t n[3], int64_t* nverts, int64_t* ncells) {
  if (dim < 1 || dim > 3) PDE_FAIL("dim must be 1, 2 or 3");
  int64_t nv = 1, nc = 1;
  for (int k = 0; k < dim; ++k) {
    if (n[k] < 1) PDE_FAIL("cell counts must be >= 1");
    nv *= (int64_t)n[k] + 1;
    nc *= n[k];
  }
  nc *= dim == 1 ? 1 : (dim == 2 ? 2 : 6);
  if (nverts) *nverts = nv;
  if (ncells) *ncells = nc;
  return 0;
}

extern "C" int pde_op_table(const pde_op_params* p, double* table, int32_t* ncomp) {
  Grid g;
  PDE_OK(make_grid(p->dim, p->n, p->L, 0, 1, &g));
  OpTable t;
  if (p->kind == PDE_OP_ELASTICITY) PDE_OK(build_elasticity_table(p->dim, g.h, g.nc, p->lam, p->mu, &t));
  else {
    double a = p->kind == PDE_OP_MASS ? 1.0 : (p->kind == PDE_OP_STIFFNESS ? 0.0 : p->alpha);
    double b = p->kind == PDE_OP_MASS ? 0.0 : (p->kind == PDE_OP_STIFFNESS ? 1.0 : p->beta);
    PDE_OK(build_scalar_table(p->dim, g.h, g.nc, a, b, &t));
  }
  if (ncomp) *ncomp = t.ncomp;
  if (table) std::memcpy(table, t.coef.data(), t.coef.size() * sizeof(double));
  return 0;
}

extern "C" int pde_slab_partition(int dim, const int32_t n_in[3], int rank, int world, int level, int32_t* z0,
                                  int32_t* nzl, int32_t* nzg) {
  if (dim < 1 || dim > 3) PDE_FAIL("dim must be 1, 2 or 3");
  if (world < 1 || rank < 0 || rank >= world) PDE_FAIL("bad rank/world");
  int32_t n[3] = {0, 0, 0};
  for (int k = 0; k < dim; ++k) n[k] = n_in[k];
  for (int l = 0; l < level; ++l) {
    for (int k = 0; k < dim; ++k) {
      if (n[k] % 2 != 0 || n[k] < 2) PDE_FAIL("level does not exist: an axis cannot be halved");
      n[k] /= 2;
    }
    if (world > 1 && (n[dim - 1] * 2) % (2 * world) != 0) PDE_FAIL("level does not exist: coarse slabs would not nest");
  }
  const double L1[3] = {1, 1, 1};
  Grid g;
  PDE_OK(make_grid(dim, n, L1, rank, world, &g));
  if (z0) *z0 = g.z0;
  if (nzl) *nzl = g.nzl;
  if (nzg) *nzg = g.nzg;
  return 0;
}
