// Byte-code interpreter for closed-form expressions of x, t, T, P1 fields and
// species concentrations (sources, fluxes, reaction rates and their species
// derivatives).  Op codes are shared with festim_b200/ufl.py (ProgramBuilder).
#pragma once
#include "fb_internal.cuh"

enum {
  OP_CONST = 0, OP_X = 1, OP_SCALAR = 2, OP_FIELD = 3, OP_FIELDGRAD = 4, OP_SPECIES = 5,
  OP_T = 6, OP_SPECIES_PREV = 7,
  OP_ADD = 10, OP_MUL = 11, OP_DIV = 12, OP_POW = 13, OP_NEG = 14,
  OP_MATH = 20,
  OP_LT = 30, OP_GT = 31, OP_LE = 32, OP_GE = 33, OP_EQ = 34, OP_NE = 35, OP_AND = 36,
  OP_OR = 37, OP_NOT = 38,
  OP_SELECT = 40, OP_MAX = 41, OP_MIN = 42
};

// Evaluation point: physical coordinates, temperature, species values and the
// entity (cell or facet) it lies in, for lazy P1-field interpolation.
struct PointCtx {
  double x[3];
  double T;
  const double* c;       // [ns] species at the point
  const double* cprev;   // [ns] previous-step species at the point
  const int* verts;      // vertices of the integration entity
  const double* bary;    // barycentric coordinates w.r.t. verts
  int nverts;
  const int* cverts;     // vertices of the owning cell
  const double* G;       // [ncverts][tdim] basis gradients of the owning cell
  int ncverts;
  int tdim;
};

__device__ __forceinline__ double fb_math(int f, double a) {
  switch (f) {
    case 0: return exp(a);
    case 1: return log(a);
    case 2: return sqrt(a);
    case 3: return sin(a);
    case 4: return cos(a);
    case 5: return tan(a);
    case 6: return tanh(a);
    case 7: return sinh(a);
    case 8: return cosh(a);
    case 9: return fabs(a);
    case 10: return erf(a);
    case 11: return atan(a);
    case 12: return asin(a);
    case 13: return acos(a);
    case 14: return (a > 0.0) - (a < 0.0);
  }
  return 0.0;
}

__device__ inline double fb_prog_eval(const DevSpec& S, int pid, const PointCtx& P) {
  double st[FB_STACK];
  int sp = 0;
  const int beg = S.prog_offsets[pid], end = S.prog_offsets[pid + 1];
  for (int pc = beg; pc < end; ++pc) {
    const int op = S.prog_ops[pc];
    const int ia = S.prog_iargs[pc];
    switch (op) {
      case OP_CONST: st[sp++] = S.prog_fargs[pc]; break;
      case OP_X: st[sp++] = P.x[ia]; break;
      case OP_SCALAR: st[sp++] = S.scalars[ia]; break;
      case OP_T: st[sp++] = P.T; break;
      case OP_SPECIES: st[sp++] = P.c[ia]; break;
      case OP_SPECIES_PREV: st[sp++] = P.cprev[ia]; break;
      case OP_FIELD: {
        const double* f = S.fields[ia];
        double v = 0.0;
        for (int a = 0; a < P.nverts; ++a) v += P.bary[a] * f[P.verts[a]];
        st[sp++] = v;
      } break;
      case OP_FIELDGRAD: {
        const double* f = S.fields[ia >> 2];
        const int comp = ia & 3;
        double v = 0.0;
        for (int a = 0; a < P.ncverts; ++a) v += f[P.cverts[a]] * P.G[a * P.tdim + comp];
        st[sp++] = v;
      } break;
      case OP_ADD: sp--; st[sp - 1] += st[sp]; break;
      case OP_MUL: sp--; st[sp - 1] *= st[sp]; break;
      case OP_DIV: sp--; st[sp - 1] /= st[sp]; break;
      case OP_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case OP_MATH: st[sp - 1] = fb_math(ia, st[sp - 1]); break;
      case OP_LT: sp--; st[sp - 1] = st[sp - 1] < st[sp]; break;
      case OP_GT: sp--; st[sp - 1] = st[sp - 1] > st[sp]; break;
      case OP_LE: sp--; st[sp - 1] = st[sp - 1] <= st[sp]; break;
      case OP_GE: sp--; st[sp - 1] = st[sp - 1] >= st[sp]; break;
      case OP_EQ: sp--; st[sp - 1] = st[sp - 1] == st[sp]; break;
      case OP_NE: sp--; st[sp - 1] = st[sp - 1] != st[sp]; break;
      case OP_AND: sp--; st[sp - 1] = (st[sp - 1] != 0.0) && (st[sp] != 0.0); break;
      case OP_OR: sp--; st[sp - 1] = (st[sp - 1] != 0.0) || (st[sp] != 0.0); break;
      case OP_NOT: st[sp - 1] = (st[sp - 1] == 0.0); break;
      case OP_SELECT: sp -= 2; st[sp - 1] = (st[sp - 1] != 0.0) ? st[sp] : st[sp + 1]; break;
      case OP_MAX: sp--; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
      case OP_MIN: sp--; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
    }
  }
  return st[0];
}
