"""Oracle: a numpy stand-in for the slice of `dolfin` the reference's hot path touches, so that the reference's OWN
time loops (helpers.py solve_schnak_system, solve_adjoint_schnak_system, solve_nonlinear_equation,
solve_adjoint_nonlinear_equation, solve_chtxs_system, solve_adjoint_chtxs_system, armijo_line_search_ref) run here
UNMODIFIED and produce golden vectors (tests/golden/make_golden.py -> ref_loops.npz).

TEST INFRASTRUCTURE (see oracle/__init__.py); used in the build container only.

What it provides (everything helpers.py names): TrialFunction, TestFunction, Function (+ .vector().set_local),
Constant, Expression (C-syntax component strings with keyword parameters, settable attributes such as wind.t),
dx, dot, grad, div, exp, assemble, as_backend_type(A).mat().getValuesCSR()/.size, and a FunctionSpace stand-in.

How a form is assembled: the integrand is kept as an expression tree; `assemble` evaluates it for every cell at the
quadrature points of FIAT's default triangle scheme of the degree UFL would estimate (sum over the factors of a product,
max over the terms of a sum; P1 argument / coefficient 1, its gradient 0, Constant 0, Expression(degree=k) k,
exp(f) deg(f)+2, f**p p*deg(f)) -- SURVEY.md App. B.3 -- and sums the element tensors cell by cell into the fixed CSR
pattern (explicit zeros kept: helpers.py:87-104).  Expression(degree=4) coefficients are evaluated at the quadrature
points directly; dolfin interpolates them to P4 per cell first, which is exact for the reference's polynomial winds
(helpers.py:506-508, 876-878), the only Expressions on the hot path.  The tree evaluator is independent of the
hand-written element tensors of oracle/p1assembly.py (which it uses for geometry, quadrature tables and the scatter), so
tests/test_oracle_golden.py also compares the two form by form.

dolfin / FFC / FIAT / PETSc themselves are third-party and absent from /root/reference (no pinned version: cpython-38
byte-code => FEniCS-legacy 2019.x); this is a restatement of their published algorithm for P1 on triangles, pinned
end to end by the reference's shipped chemotaxis and solid-body trajectories (tests/test_oracle_golden.py).
"""
import math
import re
import types

import numpy as np

from .p1assembly import P1Assembler, quad_rule


# ----------------------------------------------------------------------------------------------------------------
# expression tree
# ----------------------------------------------------------------------------------------------------------------
class Node:
    """scalar- or vector-valued integrand node; value arrays broadcast over [cell, quad point, test i, trial j(, component)]"""
    rank = 0          # 0 scalar, 1 vector

    def __add__(self, o): return Sum(self, _wrap(o))
    def __radd__(self, o): return Sum(_wrap(o), self)
    def __sub__(self, o): return Sum(self, Prod(Const(-1.0), _wrap(o)))
    def __rsub__(self, o): return Sum(_wrap(o), Prod(Const(-1.0), self))
    def __neg__(self): return Prod(Const(-1.0), self)

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([self])
        return Prod(self, _wrap(o))

    def __rmul__(self, o): return Prod(_wrap(o), self)
    def __truediv__(self, o): return Div(self, _wrap(o))
    def __rtruediv__(self, o): return Div(_wrap(o), self)
    def __pow__(self, p): return Pow(self, p)

    def args(self):
        return set()


def _wrap(o):
    if isinstance(o, Node):
        return o
    if isinstance(o, (int, float, np.floating, np.integer)):
        return Const(float(o))
    raise TypeError(f"cannot use {type(o).__name__} in a form")


class Const(Node):
    def __init__(self, v): self.v = float(v)
    def degree(self): return 0
    def eval(self, ctx): return self.v
    def grad_eval(self, ctx): return np.zeros((1, 1, 1, 1, 2))


class Argument(Node):
    """TrialFunction (number 1) / TestFunction (number 0)"""
    def __init__(self, V, number): self.V, self.number = V, number
    def degree(self): return 1
    def args(self): return {self.number}

    def eval(self, ctx):
        phi = ctx.phi                                   # [q, a]
        return phi[None, :, :, None] if self.number == 0 else phi[None, :, None, :]

    def grad_eval(self, ctx):
        G = ctx.asm.G                                   # [c, a, 2]
        return G[:, None, :, None, :] if self.number == 0 else G[:, None, None, :, :]


class Function(Node):
    """P1 coefficient (helpers.py:123-141 vec_to_function)"""
    def __init__(self, V):
        self.V = V
        self.vec = np.zeros(V.dim())

    def vector(self):
        return self

    def set_local(self, vec):
        self.vec = np.array(vec, dtype=np.float64).ravel()

    def get_local(self):
        return self.vec.copy()

    def degree(self): return 1
    def eval(self, ctx): return ctx.asm.at_quad(self.vec, ctx.phi)[:, :, None, None]

    def grad_eval(self, ctx):
        g = np.einsum('ca,cad->cd', self.vec[ctx.asm.cells], ctx.asm.G)
        return g[:, None, None, None, :]


class Expression(Node):
    """dolfin.Expression with C-syntax strings ("x[0]", "x[1]", pow, sin, cos, exp, pi + keyword parameters).
    A tuple of strings gives a vector expression."""

    def __init__(self, code, degree=None, **params):
        object.__setattr__(self, "_code", code if isinstance(code, (tuple, list)) else (code,))
        object.__setattr__(self, "rank", 1 if isinstance(code, (tuple, list)) else 0)
        object.__setattr__(self, "_degree", degree)
        object.__setattr__(self, "_params", dict(params))

    def __setattr__(self, k, v):                        # wind.t = t
        self._params[k] = v

    def __getattr__(self, k):
        p = object.__getattribute__(self, "_params")
        if k in p:
            return p[k]
        raise AttributeError(k)

    def degree(self):
        return 2 if self._degree is None else int(self._degree)

    def _component(self, s, X, Y):
        env = {"x": (X, Y), "pi": math.pi, "sin": np.sin, "cos": np.cos, "exp": np.exp, "pow": np.power,
               "sqrt": np.sqrt, "fabs": np.abs}
        env.update(self._params)
        return eval(re.sub(r"\bDOLFIN_PI\b", "pi", s), {"__builtins__": {}}, env) + 0.0 * X

    def eval(self, ctx):
        xy = ctx.xyq                                     # [c, q, 2]
        comps = [self._component(s, xy[..., 0], xy[..., 1]) for s in self._code]
        if self.rank == 0:
            return comps[0][:, :, None, None]
        return np.stack(comps, axis=-1)[:, :, None, None, :]


def Constant(v):
    if isinstance(v, (tuple, list, np.ndarray)):
        return ConstVec(v)
    return Const(v)


class ConstVec(Node):
    rank = 1
    def __init__(self, v): self.v = np.array(v, dtype=np.float64)
    def degree(self): return 0
    def eval(self, ctx): return self.v[None, None, None, None, :]


class Sum(Node):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.rank = max(a.rank, b.rank)
    def degree(self): return max(self.a.degree(), self.b.degree())
    def args(self): return self.a.args() | self.b.args()
    def eval(self, ctx): return self.a.eval(ctx) + self.b.eval(ctx)


class Prod(Node):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.rank = max(a.rank, b.rank)
        if a.rank and b.rank:
            raise TypeError("use dot() for vector * vector")
    def degree(self): return self.a.degree() + self.b.degree()
    def args(self): return self.a.args() | self.b.args()

    def grad_eval(self, ctx):            # product rule for scalar factors (needed by div(f * grad(p)))
        if self.a.rank or self.b.rank:
            raise TypeError("gradient of a vector-valued product")
        x, y = self.a.eval(ctx), self.b.eval(ctx)
        x = x[..., None] if isinstance(x, np.ndarray) else x
        y = y[..., None] if isinstance(y, np.ndarray) else y
        return x * self.b.grad_eval(ctx) + y * self.a.grad_eval(ctx)

    def eval(self, ctx):
        x, y = self.a.eval(ctx), self.b.eval(ctx)
        if self.a.rank and not self.b.rank and isinstance(y, np.ndarray):
            y = y[..., None]
        if self.b.rank and not self.a.rank and isinstance(x, np.ndarray):
            x = x[..., None]
        return x * y


class Div(Node):
    def __init__(self, a, b): self.a, self.b = a, b
    def degree(self): return self.a.degree() + self.b.degree()
    def args(self): return self.a.args() | self.b.args()
    def eval(self, ctx): return self.a.eval(ctx) / self.b.eval(ctx)


class Pow(Node):
    def __init__(self, a, p): self.a, self.p = a, p
    def degree(self):
        p = self.p
        return self.a.degree() * int(p) if float(p).is_integer() and p >= 0 else self.a.degree() + 2
    def args(self): return self.a.args()
    def eval(self, ctx): return self.a.eval(ctx) ** self.p


class Exp(Node):
    def __init__(self, a): self.a = a
    def degree(self): return self.a.degree() + 2
    def args(self): return self.a.args()
    def eval(self, ctx): return np.exp(self.a.eval(ctx))


class Grad(Node):
    rank = 1
    def __init__(self, a):
        if not hasattr(a, "grad_eval"):
            raise TypeError("grad() of a P1 function or argument only")
        self.a = a
    def degree(self): return max(self.a.degree() - 1, 0)
    def args(self): return self.a.args()
    def eval(self, ctx): return self.a.grad_eval(ctx)


class Dot(Node):
    def __init__(self, a, b): self.a, self.b = a, b
    def degree(self): return self.a.degree() + self.b.degree()
    def args(self): return self.a.args() | self.b.args()
    def eval(self, ctx): return np.sum(self.a.eval(ctx) * self.b.eval(ctx), axis=-1)


class VecFunction(Node):
    """P1 vector field, e.g. project(wind, VectorFunctionSpace(mesh, 'CG', 1)) (Schnak_FCT_PDECO.py:70,242)"""
    rank = 1
    def __init__(self, W, wx, wy):
        self.W, self.V = W, W.scalar
        self.wx, self.wy = np.array(wx, dtype=np.float64), np.array(wy, dtype=np.float64)
    def degree(self): return 1
    def eval(self, ctx):
        return np.stack([ctx.asm.at_quad(self.wx, ctx.phi), ctx.asm.at_quad(self.wy, ctx.phi)], axis=-1)[:, :, None, None, :]
    def div_eval(self, ctx):
        G, c = ctx.asm.G, ctx.asm.cells
        return (np.einsum('ca,ca->c', self.wx[c], G[:, :, 0]) + np.einsum('ca,ca->c', self.wy[c], G[:, :, 1]))[:, None, None, None]


class DivOp(Node):
    """div(w_h * u) = div(w_h) u + w_h . grad(u) for a P1 vector field and an argument (Schnak_FCT_PDECO.py:256);
    div(grad(f)) of a P1 function vanishes cell-wise (mimura_data_helpers.py:105)"""
    def __init__(self, a):
        self.kind = None
        if isinstance(a, Grad):
            self.kind = "lap"
        elif isinstance(a, Prod) and isinstance(a.b, Grad) and not a.a.rank:
            self.kind = "fgrad"          # div(f grad(p)) = grad(f) . grad(p) + f lap(p), lap(p) = 0 cell-wise for P1
        elif isinstance(a, Prod) and {type(a.a), type(a.b)} == {VecFunction, Argument}:
            self.kind = "wu"
            self.w = a.a if isinstance(a.a, VecFunction) else a.b
            self.u = a.a if isinstance(a.a, Argument) else a.b
        else:
            raise NotImplementedError("div() of a general vector field")
        self.a = a
    def degree(self): return 0 if self.kind == "lap" else max(self.a.degree() - 1, 0) if self.kind == "fgrad" else 1
    def args(self): return self.a.args()
    def eval(self, ctx):
        if self.kind == "lap":
            return 0.0
        if self.kind == "fgrad":
            return np.sum(self.a.a.grad_eval(ctx) * self.a.b.eval(ctx), axis=-1)
        return self.w.div_eval(ctx) * self.u.eval(ctx) + np.sum(self.w.eval(ctx) * self.u.grad_eval(ctx), axis=-1)


def grad(f): return Grad(f)
def dot(a, b): return Dot(_wrap(a), _wrap(b))
def exp(f): return Exp(_wrap(f))
def div(f): return DivOp(f)


class Measure:
    def __rmul__(self, integrand):
        return Form([_wrap(integrand)])


dx = Measure()


class Form:
    """sum of integrals over the domain"""
    def __init__(self, integrands): self.integrands = list(integrands)
    def __add__(self, o): return Form(self.integrands + o.integrands)
    def __sub__(self, o): return Form(self.integrands + [Prod(Const(-1.0), g) for g in o.integrands])
    def __neg__(self): return Form([Prod(Const(-1.0), g) for g in self.integrands])
    def __rmul__(self, s): return Form([Prod(_wrap(s), g) for g in self.integrands])


# ----------------------------------------------------------------------------------------------------------------
# function space, assembly
# ----------------------------------------------------------------------------------------------------------------
class FunctionSpace:
    """stand-in for dolfin.FunctionSpace(mesh, 'CG', 1) over an oracle RectMesh"""
    def __init__(self, mesh, family="CG", degree=1):
        assert family in ("CG", "P", "Lagrange") and degree == 1
        self._mesh = mesh
        self.asm = P1Assembler(mesh)
    def dim(self): return self._mesh.nodes
    def mesh(self): return self._mesh


class VectorFunctionSpace:
    """stand-in for dolfin.VectorFunctionSpace(mesh, 'CG', 1)"""
    def __init__(self, mesh, family="CG", degree=1):
        assert family in ("CG", "P", "Lagrange") and degree == 1
        self._mesh = mesh
        self.scalar = FunctionSpace(mesh)
    def mesh(self): return self._mesh
    def dim(self): return 2 * self._mesh.nodes


def project(expr, W):
    """dolfin.project(expr, W): the L2 projection, M w_k = int expr_k v dx per component (dolfin solves with LU)"""
    from scipy.sparse.linalg import spsolve
    V = W.scalar
    v = TestFunction(V)
    M = V.asm.to_csr(V.asm.mass()).tocsc()
    comps = []
    for k in range(2):
        comp = _Component(expr, k)
        comps.append(spsolve(M, assemble(comp * v * dx)))
    return VecFunction(W, comps[0], comps[1])


class _Component(Node):
    def __init__(self, vec, k): self.vec, self.k = vec, k
    def degree(self): return self.vec.degree()
    def eval(self, ctx): return self.vec.eval(ctx)[..., self.k]


def TrialFunction(V): return Argument(V, 1)
def TestFunction(V): return Argument(V, 0)


def vertex_to_dof_map(V):
    return np.array(V.mesh().vertex_to_dof)


class _Ctx:
    def __init__(self, asm, degree):
        self.asm = asm
        _, self.w, self.phi = quad_rule(max(int(degree), 1))
        self.xyq = asm.xy_quad(self.phi)


def _space_of(node):
    for attr in ("a", "b"):
        sub = getattr(node, attr, None)
        if isinstance(sub, Node):
            V = _space_of(sub)
            if V is not None:
                return V
    return getattr(node, "V", None)


class AssembledMatrix:
    """what dolfin.assemble returns for a bilinear form; as_backend_type(A).mat() mimics the PETSc Mat calls of
    helpers.py:101-103"""
    def __init__(self, asm, vals):
        self.asm, self.vals = asm, vals
        self.size = (asm.n, asm.n)
    def mat(self): return self
    def getValuesCSR(self): return (self.asm.rowptr.copy(), self.asm.colidx.copy(), self.vals.copy())
    def array(self): return self.asm.to_csr(self.vals).toarray()


def as_backend_type(A):
    return A


def assemble(form):
    """dolfin.assemble for the P1 forms of the hot path: bilinear -> AssembledMatrix (values on the full P1 pattern, explicit
    zeros kept), linear -> numpy vector"""
    if not isinstance(form, Form):
        raise TypeError("assemble() needs a form (integrand * dx)")
    out = None
    for g in form.integrands:
        V = _space_of(g)
        if V is None:
            raise ValueError("form without a function space")
        asm = V.asm
        ctx = _Ctx(asm, g.degree())
        val = g.eval(ctx)
        arity = g.args()
        wq = ctx.w[None, :, None, None]
        if arity == {0, 1}:
            full = np.broadcast_to(val, (asm.cells.shape[0], ctx.w.size, 3, 3))
            local = np.sum(full * wq, axis=1) * asm.detJ[:, None, None]
            res = asm.scatter_matrix(local)
        elif arity == {0}:
            full = np.broadcast_to(val, (asm.cells.shape[0], ctx.w.size, 3, 1))
            local = np.sum(full * wq, axis=1)[:, :, 0] * asm.detJ[:, None]
            res = asm.scatter_vector(local)
        else:
            raise NotImplementedError("only linear and bilinear forms are assembled on the hot path")
        if out is None:
            out = (arity, asm, res)
        else:
            assert out[0] == arity
            out = (arity, asm, out[2] + res)
    arity, asm, res = out
    return AssembledMatrix(asm, res) if arity == {0, 1} else res


def make_module():
    """a module object that can stand in for `dolfin` in sys.modules"""
    m = types.ModuleType("dolfin")
    for name in ("TrialFunction", "TestFunction", "Function", "Constant", "Expression", "FunctionSpace", "dx", "dot", "grad",
                 "div", "exp", "assemble", "as_backend_type", "vertex_to_dof_map", "VectorFunctionSpace", "project"):
        setattr(m, name, globals()[name])
    return m
