"""Config 4b of BASELINE.json: 3D p=4 n=192 (5.3e9 nnz, 42 GB of values) on one B200, assembled in
row chunks of the first axis so that the stage workspaces fit beside fields and output."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import _device, _lib, assemblers, bspline, geometry


def main():
    p, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 192)
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    for form, fid in (('stiffness', _lib.FORM_STIFFNESS), ('mass', _lib.FORM_MASS)):
        dev = assemblers.DeviceAssembler(kvs, None, fid)
        out = be.empty(dev.nnz)
        budget = be.free_bytes() - (4 << 30)
        chunks = dev.row_chunks(None, budget)
        ws = be.empty(max(dev.workspace_bytes(c) for c in chunks), np.uint8)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        res = []
        for it in range(3):
            e[0].record()
            dev.tabulate()
            dev.compute_fields(geo)
            e[1].record()
            dev.assemble_mlb(out=out, workspace=ws)
            e[2].record()
            torch.cuda.synchronize()
            res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        k2, k3 = res[-1]
        ws_gb = round(be.nbytes(ws) / 1e9, 1)
        del ws
        torch.cuda.empty_cache()
        # spot-check against the per-entry kernel on a sample of entries of the last row chunk
        rs = dev.row_start0()
        inner = int(np.prod(dev.nband[1:], dtype=np.int64))
        rng = np.random.default_rng(0)
        S = dev.structure
        mus = [rng.integers(0, len(b), 2000) for b in S.bidx]
        I = np.zeros(2000, dtype=np.uint64)
        J = np.zeros(2000, dtype=np.uint64)
        for k in range(3):
            I = I * np.uint64(S.bs[k][0]) + S.bidx[k][mus[k], 0].astype(np.uint64)
            J = J * np.uint64(S.bs[k][1]) + S.bidx[k][mus[k], 1].astype(np.uint64)
        want = be.to_host(dev.multi_entries_device(np.column_stack((I, J))))
        flat = (mus[0].astype(np.int64) * len(S.bidx[1]) + mus[1]) * len(S.bidx[2]) + mus[2]
        got = be.to_host(out[torch.from_numpy(flat).cuda()])
        scale = float(np.abs(want).max())
        print(json.dumps({'form': form, 'p': p, 'n': n, 'nnz': dev.nnz, 'chunks': len(chunks),
                          'ws_GB': ws_gb, 'k1k2_ms': round(k2, 2), 'k3_ms': round(k3, 2),
                          'nnz_per_s': dev.nnz / ((k2 + k3) * 1e-3),
                          'max_err_vs_entrywise_rel': float(np.abs(got - want).max() / scale)}), flush=True)
        del out, dev
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
