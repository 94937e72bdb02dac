"""Extract the columns the roofline tables use from an `ncu --set full` report (raw page as CSV on stdin or a path):
    ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv | python scripts/extract_ncu_full.py "header comment" > profiles/ncu_full_r2_kernels.csv
Output columns match profiles/ncu_full_r1_kernels.csv (bench.py reads dram_rd + dram_wr of the gate GEMM from it at run time)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def find(sub):
    if sub in col:
        return col[sub]
    for h, i in col.items():
        if h.endswith('.' + sub):
            return i
    for h, i in col.items():
        if sub in h:
            return i
    return None


want = [('dram_rd', 'dram__bytes_read.sum'), ('dram_wr', 'dram__bytes_write.sum'), ('dram_pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('ms', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'), ('regs', 'launch__registers_per_thread'),
        ('smem_dyn', 'launch__shared_mem_per_block_dynamic'), ('l2_hit', 'lts__t_sector_hit_rate.pct'),
        ('tensor_pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('utchmma_fp16_pct', 'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed'),
        ('sm_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), ('warps_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active')]
idx = [(n, find(m)) for n, m in want]
for c in sys.argv[1:]:
    print('# ' + c)
print('kernel,' + ','.join(n for n, _ in idx) + '   # units: ,' + ','.join(units[i] if i is not None else '' for _, i in idx))
for r in rows[2:]:
    name = r[col['Kernel Name']].split('(')[0]
    print(name + ',' + ','.join((r[i].replace(',', '') if i is not None else '') for _, i in idx))
