"""Per-shape ncu counters of the eight GEMM launches of one 7B NExT-QA layer, this repo's kernel next to the kernel cuBLAS picks
for the same (plain) product: time, L2 -> SM read sectors, DRAM bytes, tensor-pipe activity.

    ITERS=1 WARM=1 ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,\\
dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none --csv --log-file gpurun_out/gemm_shapes_ncu.csv python tools/gemm_step_shapes.py
    python tools/gemm_shapes_ncu.py gpurun_out/gemm_shapes_ncu.csv > profiles/r2_gemm_shapes_ncu.txt

tools/gemm_step_shapes.py launches, per case, (WARM + ITERS) x this repo's kernel and then (WARM + ITERS) x cuBLAS; only GEMM-like
kernels are kept and the LAST launch of every group is reported."""
import csv
import sys

CASES = ["qkv+rope      N=12288 K=4096 ", "wo f32+res    N=4096  K=4096 ", "w13+swiglu    N=22016 K=4096 ", "w2 f32+res    N=4096  K=11008",
         "w2t+swiglu'   N=11008 K=4096 ", "w13t dX       N=4096  K=22016", "wot dX        N=4096  K=4096 ", "wqkvt dX      N=4096  K=12288"]
NN = ["w2+swiglu' NN N=11008 K=4096 ", "w13 dX NN     N=4096  K=22016", "wo dX NN      N=4096  K=4096 ", "wqkv dX NN    N=4096  K=12288"]


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    order = []
    for r in rows[1:]:
        k = int(r[ii])
        if k not in launches:
            launches[k] = {"name": r[ki]}
            order.append(k)
        launches[k][r[mi]] = float(r[vi].replace(",", ""))
    # tools/gemm_step_shapes.py puts one fill kernel (MARKER.fill_) in front of every timed case: split there, keep the LONGEST
    # launch of each segment (the GEMM itself; cuBLAS may add a small split-K reduce or a copy)
    groups, cur, started = [], [], False
    for k in order:
        l = launches[k]
        if "FillFunctor" in l["name"]:
            if started and cur:
                groups.append(cur)
            cur, started = [], True
        elif started:
            cur.append(l)
    if cur:
        groups.append(cur)
    groups = [[max(g, key=lambda l: l.get("gpu__time_duration.sum", 0.0))] for g in groups if g]
    groups = [g for g in groups if any(t in g[0]["name"] for t in ("gemm", "nvjet", "cutlass", "xmma"))]    # set-up segments hold no GEMM
    def fmt(g):
        l = g[-1]
        t = l.get("gpu__time_duration.sum", 0.0) / 1e3
        return (f"{t:7.1f} us  L2 rd {l.get('lts__t_sectors_op_read.sum', 0) * 32 / 1e6:7.0f} MB  (from SMs {l.get('lts__t_sectors_srcunit_tex_op_read.sum', 0) * 32 / 1e6:7.0f} MB)"
                f"  DRAM {(l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)) / 1e6:6.0f} MB"
                f"  tensor {l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f} %  {l['name'][:60]}")
    names = [(c, 2) for c in CASES] + [(c, 1) for c in NN]
    gi = 0
    for name, n in names:
        for j in range(n):
            if gi < len(groups):
                print(f"{name} {'this repo' if j == 0 else 'cuBLAS   '} {fmt(groups[gi])}")
                gi += 1
    if gi != len(groups):
        print(f"# {len(groups) - gi} unmatched kernel groups (cuBLAS split a product into several kernels?):")
        for g in groups[gi:]:
            print("#   " + fmt(g))


if __name__ == "__main__":
    main()
