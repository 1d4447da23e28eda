import torch, time
import numpy as np, scipy.sparse as sp
n = 200000
rng = np.random.default_rng(0)
offs = [-4000, -31, -7, -2, -1, 0, 1, 3, 9, 40, 5000]
A = sp.diags([rng.random(n - abs(o)) for o in offs], offs, shape=(n, n), format="csr")
rows = np.repeat(np.arange(n), 4)
cols = np.clip(rows // 4 + np.tile(np.array([-1, 0, 1, 2]), n), 0, n // 4 - 1)
P = sp.csr_matrix((rng.random(4 * n), (rows, cols)), shape=(n, n // 4))
P.sum_duplicates()
def tocsr(M, idt):
    return torch.sparse_csr_tensor(torch.from_numpy(M.indptr.astype(idt)).cuda(), torch.from_numpy(M.indices.astype(idt)).cuda(),
                                   torch.from_numpy(M.data).cuda(), size=M.shape)
for idt in (np.int32, np.int64):
    try:
        At, Pt, Rt = tocsr(A, idt), tocsr(P, idt), tocsr(P.T.tocsr(), idt)
        torch.cuda.synchronize(); t = time.time()
        AP = torch.sparse.mm(At, Pt)
        Ac = torch.sparse.mm(Rt, AP)
        torch.cuda.synchronize(); dt = time.time() - t
        ref = (P.T @ A @ P).tocsr(); ref.sort_indices()
        crow, col, val = Ac.crow_indices().cpu().numpy(), Ac.col_indices().cpu().numpy(), Ac.values().cpu().numpy()
        got = sp.csr_matrix((val, col, crow), shape=ref.shape)
        sorted_ok = got.has_sorted_indices
        got.sort_indices()
        print(idt.__name__, "layout", Ac.layout, "dtype idx", Ac.crow_indices().dtype, "time %.3f" % dt, "nnz", got.nnz, ref.nnz,
              "sorted", sorted_ok, "err", abs(got - ref).max())
    except Exception as e:
        print(idt.__name__, "FAILED", repr(e)[:300])
