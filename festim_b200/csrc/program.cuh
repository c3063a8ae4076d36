// Byte-code interpreter for closed-form expressions of x, t, T, P1 fields and
// species concentrations (sources, fluxes, reaction rates and their species
// derivatives).  Op codes are shared with festim_b200/ufl.py (ProgramBuilder).
#pragma once
#include "fb_device.cuh"

enum {
  OP_CONST = 0, OP_X = 1, OP_SCALAR = 2, OP_FIELD = 3, OP_FIELDGRAD = 4, OP_SPECIES = 5,
  OP_T = 6, OP_SPECIES_PREV = 7, OP_SPECIES_GRAD = 8, OP_NORMAL = 9, OP_CIRCUMRADIUS = 15,
  OP_ADD = 10, OP_MUL = 11, OP_DIV = 12, OP_POW = 13, OP_NEG = 14,
  OP_MATH = 20,
  OP_LT = 30, OP_GT = 31, OP_LE = 32, OP_GE = 33, OP_EQ = 34, OP_NE = 35, OP_AND = 36,
  OP_OR = 37, OP_NOT = 38,
  OP_SELECT = 40, OP_MAX = 41, OP_MIN = 42,
  OP_STORE = 50, OP_LOAD = 51
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

// Stack machine with the top of the stack kept in a register: a push spills the old
// top to local memory, a binary op reads one operand back.
__device__ inline double fb_prog_eval(const DevSpec& S, int pid, const PointCtx& P) {
  double st[FB_STACK];
  double tmp[16];
  double t = 0.0;  // top of stack
  int sp = 0;      // entries below the top that live in st[]
  const int beg = S.prog_offsets[pid], end = S.prog_offsets[pid + 1];
#define FB_PUSH(v) do { st[sp++] = t; t = (v); } while (0)
  for (int pc = beg; pc < end; ++pc) {
    const int op = S.prog_ops[pc];
    const int ia = S.prog_iargs[pc];
    switch (op) {
      case OP_CONST: FB_PUSH(S.prog_fargs[pc]); break;
      case OP_X: FB_PUSH(P.x[ia]); break;
      case OP_SCALAR: FB_PUSH(S.scalars[ia]); break;
      case OP_T: FB_PUSH(P.T); break;
      case OP_SPECIES: FB_PUSH(P.c[ia]); break;
      case OP_SPECIES_PREV: FB_PUSH(P.cprev[ia]); break;
      case OP_SPECIES_GRAD: FB_PUSH(P.cgrad ? P.cgrad[(ia >> 2) * 3 + (ia & 3)] : 0.0); break;
      case OP_NORMAL: FB_PUSH(P.n[ia]); break;
      case OP_CIRCUMRADIUS: FB_PUSH(P.h); break;
      case OP_LOAD: FB_PUSH(tmp[ia]); break;
      case OP_STORE: tmp[ia] = t; break;
      case OP_FIELD: {
        const double* f = S.fields[ia];
        double v = 0.0;
        for (int a = 0; a < P.nverts; ++a) v += P.bary[a] * f[P.verts[a]];
        FB_PUSH(v);
      } break;
      case OP_FIELDGRAD: {
        const double* f = S.fields[ia >> 2];
        const int comp = ia & 3;
        double v = 0.0;
        for (int a = 0; a < P.ncverts; ++a) v += f[P.cverts[a]] * P.G[a * P.tdim + comp];
        FB_PUSH(v);
      } break;
      case OP_ADD: t = st[--sp] + t; break;
      case OP_MUL: t = st[--sp] * t; break;
      case OP_DIV: t = st[--sp] / t; break;
      case OP_POW: t = pow(st[--sp], t); break;
      case OP_NEG: t = -t; break;
      case OP_MATH: t = fb_math(ia, t); break;
      case OP_LT: t = st[--sp] < t; break;
      case OP_GT: t = st[--sp] > t; break;
      case OP_LE: t = st[--sp] <= t; break;
      case OP_GE: t = st[--sp] >= t; break;
      case OP_EQ: t = st[--sp] == t; break;
      case OP_NE: t = st[--sp] != t; break;
      case OP_AND: t = (st[--sp] != 0.0) && (t != 0.0); break;
      case OP_OR: t = (st[--sp] != 0.0) || (t != 0.0); break;
      case OP_NOT: t = (t == 0.0); break;
      case OP_SELECT: { const double b = t, a = st[--sp], c = st[--sp]; t = (c != 0.0) ? a : b; } break;
      case OP_MAX: t = fmax(st[--sp], t); break;
      case OP_MIN: t = fmin(st[--sp], t); break;
    }
  }
#undef FB_PUSH
  return t;
}
