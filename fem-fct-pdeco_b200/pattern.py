"""Host-side (numpy) logic on the fixed CSR pattern: re-embedding scipy matrices, scipy views."""
import numpy as np
import scipy.sparse as sp


class HostPattern:
    def __init__(self, rowptr, colidx):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.n = self.rowptr.size - 1
        self.nnz = int(self.rowptr[-1])
        if self.colidx.size != self.nnz:
            raise ValueError("colidx length does not match rowptr[-1]")
        self._rows = None
        self._keys = None

    @property
    def rows(self):
        if self._rows is None:
            self._rows = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rowptr))
        return self._rows

    def embed(self, mat):
        """Values of a scipy sparse (or dense) matrix on the fixed pattern.

        scipy's +, -, * prune exact zeros, so matrices reaching FCT_alg_ref may have a
        sub-pattern; the reference is immune because it indexes M[i,j], D[i,j] by dof_neighbors
        (helpers.py:1818-1822).  Entries outside the pattern must be zero."""
        if sp.issparse(mat):
            csr = mat.tocsr() if not sp.isspmatrix_csr(mat) else mat
            if csr.shape != (self.n, self.n):
                raise ValueError(f"matrix shape {csr.shape} does not match the pattern ({self.n})")
            if (csr.nnz == self.nnz and csr.has_sorted_indices and np.array_equal(csr.indptr, self.rowptr)
                    and np.array_equal(csr.indices, self.colidx)):
                return np.ascontiguousarray(csr.data, dtype=np.float64)
            coo = csr.tocoo()
            r, c, d = coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float64)
        else:
            dense = np.asarray(mat, dtype=np.float64)
            if dense.shape != (self.n, self.n):
                raise ValueError(f"matrix shape {dense.shape} does not match the pattern ({self.n})")
            r, c = np.nonzero(dense)
            d = dense[r, c]
            r = r.astype(np.int64)
            c = c.astype(np.int64)
        if self._keys is None:
            self._keys = self.rows.astype(np.int64) * self.n + self.colidx.astype(np.int64)
        key = r * self.n + c
        pos = np.searchsorted(self._keys, key)
        ok = pos < self.nnz
        ok[ok] = self._keys[pos[ok]] == key[ok]
        if not ok.all() and np.any(d[~ok] != 0.0):
            raise ValueError("matrix has nonzero entries outside the fixed P1 pattern")
        return np.bincount(pos[ok], weights=d[ok], minlength=self.nnz)

    def to_scipy(self, vals):
        return sp.csr_matrix((np.array(vals, dtype=np.float64), self.colidx.copy(), self.rowptr.copy()),
                             shape=(self.n, self.n))
