"""A small UFL-like front end so that the reference's `assemble_sparse(form)` / `assemble(form)` call sites keep
working without dolfin, with the assembly done by the GPU kernels of libfctpdeco.

    from fem_fct_pdeco_b200.forms import *      # TrialFunction, TestFunction, Constant, Expression, dx, dot, grad, exp
    u, v = TrialFunction(V), TestFunction(V)
    M  = assemble_sparse_lil(u * v * dx)                                  # helpers.py:553
    Ad = assemble_sparse(dot(grad(u), grad(v)) * dx)                      # helpers.py:555
    A  = assemble_sparse(dot(wind, grad(v)) * u * dx)                     # advection_solidbody_FCT.py:106
    b  = np.asarray(assemble((uhat_fun - u_fun) * v * dx))                # advection_solidbody_FCT_PDECO_alltime.py:255

Supported: exactly the form catalogue of SURVEY.md App. C (every `assemble` site on the reference's hot path):
P1-weighted masses, stiffness, convection with constant / polynomial (degree <= 3) / P1-nodal winds in conservative
and non-conservative form, the drift-control forms, the chemotaxis forms with `exp` coefficients, and load vectors
of P1 products, constants, `p (b.grad u) v` and `chi u exp(-eta u) grad p.grad w`.  Anything else raises
NotImplementedError -- there is no generic (CPU) fallback assembler.

The integrand is expanded into a sum of products; each product is matched against the kernel catalogue
(include/fctpdeco.h FCT_FORM_* / FCT_LOAD_*) and accumulated on the device.
"""
import numpy as np
import scipy.sparse as sp

from . import _lib

__all__ = ["TrialFunction", "TestFunction", "Function", "Constant", "Expression", "dx", "dot", "grad", "exp", "div",
           "assemble", "assemble_sparse", "assemble_sparse_lil", "vec_to_function", "VectorFunctionSpace", "project"]


# ---- expression tree -------------------------------------------------------------------------------------
class Expr:
    def __add__(self, o): return Sum([self, _wrap(o)])
    def __radd__(self, o): return Sum([_wrap(o), self])
    def __sub__(self, o): return Sum([self, Prod([Num(-1.0), _wrap(o)])])
    def __rsub__(self, o): return Sum([_wrap(o), Prod([Num(-1.0), self])])
    def __neg__(self): return Prod([Num(-1.0), self])
    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form(self)
        return Prod([self, _wrap(o)])
    def __rmul__(self, o): return Prod([_wrap(o), self])
    def __truediv__(self, o):
        o = _wrap(o)
        if not isinstance(o, Num):
            raise NotImplementedError("division by a non-constant expression")
        return Prod([Num(1.0 / o.v), self])
    def __pow__(self, k):
        if not (isinstance(k, int) and k >= 1):
            raise NotImplementedError("only positive integer powers")
        return Prod([self] * k)


class Num(Expr):
    def __init__(self, v): self.v = float(v)


class Sum(Expr):
    def __init__(self, terms): self.terms = terms


class Prod(Expr):
    def __init__(self, factors): self.factors = factors


class Arg(Expr):
    def __init__(self, V, kind): self.V, self.kind = V, kind          # kind: 'u' (trial) / 'v' (test)


class Function(Expr):
    """P1 coefficient (what helpers.py:123-141 vec_to_function returns)"""
    def __init__(self, V, vec=None):
        self.V = V
        self.vec = np.zeros(V.dim()) if vec is None else np.ascontiguousarray(vec, dtype=np.float64)

    def vector(self):
        return self

    def set_local(self, vec):            # dolfin idiom used by vec_to_function
        self.vec = np.ascontiguousarray(vec, dtype=np.float64)


class VecConst(Expr):
    def __init__(self, bx, by): self.b = (float(bx), float(by))


class PolyWind(Expr):
    """vector field whose components are polynomials of degree <= 3 in (x, y): 2 x 10 monomial coefficients"""
    def __init__(self, coefs): self.coefs = np.asarray(coefs, dtype=np.float64).reshape(2, 10)


class VecFunction(Expr):
    """P1 vector field (what dolfin's project(wind, VectorFunctionSpace(mesh, 'CG', 1)) returns): two nodal arrays"""
    def __init__(self, W, wx, wy):
        self.W = W
        self.wx = np.ascontiguousarray(wx, dtype=np.float64)
        self.wy = np.ascontiguousarray(wy, dtype=np.float64)


class DivWU(Expr):
    """div(w_h * u) with a P1 vector field w_h and the trial function u (Schnak_FCT_PDECO.py:256)"""
    def __init__(self, w, u): self.w, self.u = w, u


class Grad(Expr):
    def __init__(self, f): self.f = f


class Dot(Expr):
    def __init__(self, a, b): self.a, self.b = a, b


class Exp(Expr):
    def __init__(self, scale, f): self.scale, self.f = float(scale), f      # exp(scale * f), f a Function


class Measure:
    pass


dx = Measure()


class Form:
    def __init__(self, integrand): self.integrand = integrand
    def __add__(self, o): return Form(Sum([self.integrand, o.integrand]))
    def __sub__(self, o): return Form(Sum([self.integrand, Prod([Num(-1.0), o.integrand])]))
    def __rmul__(self, c): return Form(Prod([_wrap(c), self.integrand]))
    def __neg__(self): return Form(Prod([Num(-1.0), self.integrand]))


def _wrap(o):
    if isinstance(o, Expr):
        return o
    if isinstance(o, (int, float, np.floating, np.integer)):
        return Num(o)
    raise TypeError(f"cannot use {type(o).__name__} in a form")


def TrialFunction(V): return Arg(V, "u")
def TestFunction(V): return Arg(V, "v")


def Constant(value):
    if np.ndim(value) == 0:
        return Num(float(value))
    v = [float(x) for x in value]
    if len(v) != 2:
        raise NotImplementedError("vector constants must have two components")
    return VecConst(*v)


class _P:
    """polynomial in x, y used to evaluate Expression strings"""
    def __init__(self, c=None): self.c = dict(c or {})
    @staticmethod
    def of(o): return o if isinstance(o, _P) else _P({(0, 0): float(o)})
    def __add__(self, o):
        o = _P.of(o); r = dict(self.c)
        for k, v in o.c.items(): r[k] = r.get(k, 0.0) + v
        return _P(r)
    __radd__ = __add__
    def __neg__(self): return _P({k: -v for k, v in self.c.items()})
    def __sub__(self, o): return self + (-_P.of(o))
    def __rsub__(self, o): return _P.of(o) + (-self)
    def __mul__(self, o):
        o = _P.of(o); r = {}
        for (a, b), v in self.c.items():
            for (c, d), w in o.c.items():
                r[(a + c, b + d)] = r.get((a + c, b + d), 0.0) + v * w
        return _P(r)
    __rmul__ = __mul__
    def __truediv__(self, o): return self * (1.0 / float(o))
    def coefs(self):
        order = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2), (3, 0), (2, 1), (1, 2), (0, 3)]
        out = np.zeros(10)
        for k, v in self.c.items():
            if v == 0.0:
                continue
            if k not in order:
                raise NotImplementedError("Expression winds must be polynomials of degree <= 3")
            out[order.index(k)] = v
        return out


class Expression(PolyWind):
    """dolfin.Expression for polynomial vector fields, e.g. Expression(('-x[1]','x[0]'), degree=4) or
    ("speed*2*(x[1]-0.5)*x[0]*(1-x[0])", ...) (helpers.py:876-878).  Parsed by evaluating the C-like strings on
    polynomial objects; parameters are keywords and stay settable (`wind.t = t`, helpers.py:566), the coefficients being
    re-derived on every change."""

    def __init__(self, code, degree=None, **params):
        if isinstance(code, str) or len(code) != 2:
            raise NotImplementedError("only two-component vector Expressions are supported")
        object.__setattr__(self, "_code", tuple(str(c) for c in code))
        object.__setattr__(self, "_params", {k: float(v) for k, v in params.items()})
        self._reparse()

    def _reparse(self):
        import math
        env = {"x": [_P({(1, 0): 1.0}), _P({(0, 1): 1.0})], "__builtins__": {}, "pi": math.pi, "sin": math.sin,
               "cos": math.cos, "exp": math.exp, "sqrt": math.sqrt, "pow": pow}      # functions of the parameters only
        env.update(self._params)
        rows = [_P.of(eval(comp, env)).coefs() for comp in self._code]          # noqa: S307  (polynomial objects only)
        object.__setattr__(self, "coefs", np.array(rows).reshape(2, 10))

    def __setattr__(self, k, v):
        self._params[k] = float(v)
        self._reparse()

    def __getattr__(self, k):
        p = object.__getattribute__(self, "_params")
        if k in p:
            return p[k]
        raise AttributeError(k)


def grad(f):
    if not isinstance(f, (Arg, Function)):
        raise NotImplementedError("grad() of a trial/test/P1 function only")
    return Grad(f)


def dot(a, b):
    return Dot(a, b)


def div(a):
    """div(w_h * u) with w_h = project(wind, W) (Schnak_FCT_PDECO.py:256) is the one div form on the legacy hot path;
    div(grad f) of a P1 field vanishes cell-wise (mimura_data_helpers.py:105) and contributes nothing"""
    if isinstance(a, Prod) and len(a.factors) == 2:
        w = [f for f in a.factors if isinstance(f, VecFunction)]
        u = [f for f in a.factors if isinstance(f, Arg) and f.kind == "u"]
        if len(w) == 1 and len(u) == 1:
            return DivWU(w[0], u[0])
    if isinstance(a, Grad) and isinstance(a.f, Function):
        return Num(0.0)
    raise NotImplementedError("div(...) of this expression is not on the catalogue")


class VectorFunctionSpace:
    """stand-in for dolfin.VectorFunctionSpace(mesh, 'CG', 1)"""
    def __init__(self, mesh, family="CG", degree=1):
        if family not in ("CG", "P", "Lagrange") or degree != 1:
            raise NotImplementedError("only P1 vector spaces")
        self._mesh = mesh
    def mesh(self): return self._mesh
    def dim(self): return 2 * self._mesh.nodes


def project(wind, W):
    """dolfin.project(wind, W) for a polynomial wind (Schnak_FCT_PDECO.py:70,242): the L2 projection onto P1, component by
    component -- M w_k = int wind_k v dx, right-hand sides by the exact degree-5 rule (FCT_LOAD_POLY3), mass-matrix solves by
    Jacobi-PCG on the device (dolfin: LU)"""
    if not isinstance(wind, PolyWind):
        raise NotImplementedError("project() of a polynomial Expression only")
    ctx = W.mesh().context()
    M = ctx.static()[0]
    comps = []
    for k in range(2):
        dco = ctx.array(np.ascontiguousarray(wind.coefs[k]))
        rhs, x = ctx.empty(ctx.n), ctx.empty(ctx.n)
        ctx.assemble_vector(_lib.LOAD_POLY3, rhs, c0=dco)
        ctx.axpby(0.0, rhs, 0.0, None, x)
        ctx.solve(_lib.SOLVER_PCG, M, rhs, x, rtol=1e-14, maxit=2000)
        comps.append(x.download())
        for a in (dco, rhs, x):
            a.free()
    return VecFunction(W, comps[0], comps[1])


def exp(f):
    """exp(scale * Function): recognises exp(-eta*m) as the reference writes it"""
    scale, fn = 1.0, f
    if isinstance(f, Prod):
        nums = [x.v for x in f.factors if isinstance(x, Num)]
        fns = [x for x in f.factors if isinstance(x, Function)]
        if len(nums) + len(fns) != len(f.factors) or len(fns) != 1:
            raise NotImplementedError("exp() of a constant multiple of one P1 function only")
        scale, fn = float(np.prod(nums)) if nums else 1.0, fns[0]
    if not isinstance(fn, Function):
        raise NotImplementedError("exp() of a constant multiple of one P1 function only")
    return Exp(scale, fn)


def vec_to_function(vec, V):
    """helpers.py:123-141"""
    out = Function(V)
    out.vector().set_local(vec)
    return out


# ---- expansion into a sum of products ------------------------------------------------------------------------
def _expand(e):
    """-> list of (coef, [atoms]); atoms: Arg, Function, Exp, Dot, VecConst/PolyWind only inside Dot"""
    if isinstance(e, Num):
        return [(e.v, [])]
    if isinstance(e, Sum):
        out = []
        for t in e.terms:
            out += _expand(t)
        return out
    if isinstance(e, Prod):
        acc = [(1.0, [])]
        for f in e.factors:
            nxt = []
            for c1, a1 in acc:
                for c2, a2 in _expand(f):
                    nxt.append((c1 * c2, a1 + a2))
            acc = nxt
        return acc
    if isinstance(e, Dot):
        # distribute sums inside the arguments of dot (e.g. dot(1/om*wind + move, grad(v)))
        out = []
        for ca, xa in _expand_vec(e.a):
            for cb, xb in _expand_vec(e.b):
                out.append((ca * cb, [Dot(xa, xb)]))
        return out
    if isinstance(e, (Arg, Function, Exp, DivWU)):
        return [(1.0, [e])]
    raise NotImplementedError(f"unsupported expression node {type(e).__name__}")


def _expand_vec(e):
    if isinstance(e, (Grad, VecConst, PolyWind)):
        return [(1.0, e)]
    if isinstance(e, Sum):
        out = []
        for t in e.terms:
            out += _expand_vec(t)
        return out
    if isinstance(e, Prod):
        c, vec = 1.0, None
        for f in e.factors:
            if isinstance(f, Num):
                c *= f.v
            elif vec is None:
                vec = f
            else:
                raise NotImplementedError("product of two vector fields")
        return [(c * cc, v) for cc, v in _expand_vec(vec)]
    raise NotImplementedError(f"unsupported vector expression {type(e).__name__}")


def _merge_winds(terms):
    """dot(w1, grad(v))*u + dot(w2, grad(v))*u with polynomial / constant winds -> one polynomial wind"""
    return terms


# ---- matching against the kernel catalogue -------------------------------------------------------------------
class _Term:
    def __init__(self, coef, atoms):
        self.coef = coef
        self.u = [a for a in atoms if isinstance(a, Arg) and a.kind == "u"]
        self.v = [a for a in atoms if isinstance(a, Arg) and a.kind == "v"]
        self.fns = [a for a in atoms if isinstance(a, Function)]
        self.exps = [a for a in atoms if isinstance(a, Exp)]
        self.dots = [a for a in atoms if isinstance(a, Dot)]
        self.divs = [a for a in atoms if isinstance(a, DivWU)]
        self.u += [a.u for a in self.divs]

    def space(self):
        for a in self.u + self.v:
            return a.V
        for d in self.dots:
            for x in (d.a, d.b):
                if isinstance(x, Grad):
                    return x.f.V
        raise NotImplementedError("form without trial/test function")


def _is_grad_of(x, kind):
    return isinstance(x, Grad) and isinstance(x.f, Arg) and x.f.kind == kind


def _grad_fn(x):
    return x.f if isinstance(x, Grad) and isinstance(x.f, Function) else None


def _wind_coefs(w):
    if isinstance(w, PolyWind):
        return w.coefs
    c = np.zeros((2, 10))
    c[0, 0], c[1, 0] = w.b
    return c


def _assemble_matrix_terms(ctx, terms):
    L = _lib
    out = ctx.empty(ctx.nnz)
    first = [True]

    def emit(kind, scale, **kw):
        dev = {k: (ctx.array(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        ctx.assemble_matrix(kind, out, scale=scale, accumulate=not first[0], **dev)
        first[0] = False

    used = [False] * len(terms)
    for i, t in enumerate(terms):
        if used[i]:
            continue
        used[i] = True
        nu, nv, nf, ne, nd = len(t.u), len(t.v), len(t.fns), len(t.exps), len(t.dots)
        if len(t.divs) == 1 and nd == 0 and nu == 1 and nv == 1 and nf == 0 and ne == 0:
            # div(w_h u) v = (w_h . grad u) v + div(w_h) u v        (Schnak_FCT_PDECO.py:256)
            w = t.divs[0].w
            emit(L.FORM_WIND_P1_T, t.coef, c0=w.wx, c1=w.wy)
            emit(L.FORM_DIVW_MASS, t.coef, c0=w.wx, c1=w.wy)
            continue
        if len(t.divs):
            raise NotImplementedError("div(w_h u) outside div(w_h u) * v * dx")
        if nd == 0 and nu == 1 and nv == 1 and ne == 0:                     # (f1 f2 f3) u v
            if nf == 0:
                emit(L.FORM_MASS, t.coef)
            elif nf <= 3:
                kw = {f"c{k}": t.fns[k].vec for k in range(nf)}
                emit([L.FORM_WMASS1, L.FORM_WMASS2, L.FORM_WMASS3][nf - 1], t.coef, **kw)
            else:
                raise NotImplementedError("weighted mass with more than three P1 coefficients")
            continue
        if nd == 1 and nu == 0 and nv == 0 and nf == 0 and ne == 0:
            d = t.dots[0]
            if (_is_grad_of(d.a, "u") and _is_grad_of(d.b, "v")) or (_is_grad_of(d.a, "v") and _is_grad_of(d.b, "u")):
                emit(L.FORM_STIFFNESS, t.coef)                              # grad u . grad v
                continue
        if nd == 1:
            d = t.dots[0]
            a, b = (d.a, d.b) if isinstance(d.b, Grad) else (d.b, d.a)       # b: the gradient factor
            # (w . grad v) u   /   (w . grad u) v     with constant or polynomial wind
            if isinstance(a, (VecConst, PolyWind)) and nf == 0 and ne == 0:
                if _is_grad_of(b, "v") and nu == 1 and nv == 0:
                    emit(L.FORM_WIND_POLY3, t.coef, c0=_wind_coefs(a).ravel().copy())
                    continue
                if _is_grad_of(b, "u") and nv == 1 and nu == 0:
                    emit(L.FORM_WIND_POLY3_T, t.coef, c0=_wind_coefs(a).ravel().copy())
                    continue
            # drift-control forms: (b . grad c) u v   and   (b . grad v) c u
            if isinstance(a, VecConst) and ne == 0:
                if _grad_fn(b) is not None and nu == 1 and nv == 1 and nf == 0:
                    emit(L.FORM_DRIFT_MASS, t.coef, c0=_grad_fn(b).vec, s0=a.b[0], s1=a.b[1])
                    continue
                if _is_grad_of(b, "v") and nu == 1 and nv == 0 and nf == 1:
                    emit(L.FORM_DRIFT_CONV, t.coef, c0=t.fns[0].vec, s0=a.b[0], s1=a.b[1])
                    continue
            # chemotaxis: [exp(-eta m)] (grad f . grad v) u
            fg = _grad_fn(a) or _grad_fn(b)
            other = b if _grad_fn(a) is not None else a
            if fg is not None and _is_grad_of(other, "v") and nu == 1 and nv == 0 and nf == 0:
                if ne == 0:
                    emit(L.FORM_CHTX, t.coef, c0=fg.vec)
                    continue
                if ne == 1:
                    emit(L.FORM_CHTX_EXP, t.coef, c0=fg.vec, c1=t.exps[0].f.vec, s0=-t.exps[0].scale)
                    continue
            # adjoint chemotaxis: (1 - eta u) exp(-eta u) (grad u_trial . grad vn) w  -- arrives as two terms
            if fg is not None and _is_grad_of(other, "u") and nv == 1 and nu == 0 and ne == 1 and nf == 0:
                ex = t.exps[0]
                eta = -ex.scale
                for j in range(i + 1, len(terms)):
                    s = terms[j]
                    if used[j] or len(s.dots) != 1 or len(s.exps) != 1 or len(s.fns) != 1:
                        continue
                    sd = s.dots[0]
                    sfg = _grad_fn(sd.a) or _grad_fn(sd.b)
                    if (sfg is fg and s.exps[0].f is ex.f and s.exps[0].scale == ex.scale and s.fns[0] is ex.f
                            and abs(s.coef + eta * t.coef) <= 1e-14 * abs(t.coef)):
                        used[j] = True
                        emit(L.FORM_CHTX_ADJ, t.coef, c0=fg.vec, c1=ex.f.vec, s0=eta)
                        break
                else:
                    raise NotImplementedError("exp-weighted (grad u . grad f) w without its (1 - eta u) partner")
                continue
        raise NotImplementedError("bilinear form outside the catalogue of SURVEY.md App. C")
    return out


def _assemble_vector_terms(ctx, terms):
    L = _lib
    out = ctx.empty(ctx.n)
    first = [True]

    def emit(kind, scale, **kw):
        dev = {k: (ctx.array(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        ctx.assemble_vector(kind, out, scale=scale, accumulate=not first[0], **dev)
        first[0] = False

    for t in terms:
        nf, ne, nd = len(t.fns), len(t.exps), len(t.dots)
        if len(t.u) != 0 or len(t.v) > 1:
            raise NotImplementedError("not a linear form")
        if nd == 0 and ne == 0 and len(t.v) == 1:
            if nf == 0:
                emit(L.LOAD_CONST, t.coef, s0=1.0)
            elif nf <= 4:
                kw = {f"c{k}": t.fns[k].vec for k in range(nf)}
                emit([L.LOAD_P1_1, L.LOAD_P1_2, L.LOAD_P1_3, L.LOAD_P1_4][nf - 1], t.coef, **kw)
            else:
                raise NotImplementedError("load vector with more than four P1 factors")
            continue
        if nd == 1:
            d = t.dots[0]
            # p (b . grad u) v
            a, b = (d.a, d.b) if isinstance(d.b, Grad) else (d.b, d.a)
            if isinstance(a, VecConst) and _grad_fn(b) is not None and nf == 1 and ne == 0 and len(t.v) == 1:
                emit(L.LOAD_DRIFT_GRAD, t.coef, c0=t.fns[0].vec, c1=_grad_fn(b).vec, s0=a.b[0], s1=a.b[1])
                continue
            # chi u exp(-eta u) grad p . grad w
            fg = _grad_fn(d.a) or _grad_fn(d.b)
            other = d.b if _grad_fn(d.a) is not None else d.a
            if (fg is not None and _is_grad_of(other, "v") and len(t.v) == 0 and nf == 1 and ne == 1
                    and t.exps[0].f is t.fns[0]):
                emit(L.LOAD_CHTX_ADJ, 1.0, c0=fg.vec, c1=t.fns[0].vec, s0=-t.exps[0].scale, s1=t.coef)
                continue
        raise NotImplementedError("linear form outside the catalogue of SURVEY.md App. C")
    return out


def _terms(form):
    if not isinstance(form, Form):
        raise TypeError("expected a form (integrand * dx)")
    return [_Term(c, a) for c, a in _expand(form.integrand) if c != 0.0]


def assemble(form):
    """dolfin.assemble for linear forms: numpy vector in DoF order (np.asarray(assemble(...)) at the call sites)"""
    terms = _terms(form)
    if any(len(t.u) for t in terms) or any(_is_grad_of(x, "u") for t in terms for d in t.dots for x in (d.a, d.b)):
        raise NotImplementedError("assemble() of a bilinear form: use assemble_sparse()")
    ctx = terms[0].space().mesh().context()
    out = _assemble_vector_terms(ctx, terms)
    res = out.download()
    out.free()
    return res


def assemble_sparse(a):
    """helpers.py:87-104: scipy CSR on the full P1 pattern, explicit zeros kept, columns ascending"""
    terms = _terms(a)
    ctx = terms[0].space().mesh().context()
    out = _assemble_matrix_terms(ctx, terms)
    vals = out.download()
    out.free()
    return sp.csr_matrix((vals, ctx.colidx.copy(), ctx.rowptr.copy()), shape=(ctx.n, ctx.n))


def assemble_sparse_lil(a):
    """helpers.py:106-121"""
    return sp.lil_matrix(assemble_sparse(a))
