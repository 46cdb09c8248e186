"""Select the judged metrics from `ncu -i rep --page raw --csv` output into a compact per-launch table.
usage: ncu_extract.py <raw.csv> [stride] > profiles/<name>.csv   (stride 2 keeps every second launch: the warm one)"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max.per_second']
rows = list(csv.reader(open(sys.argv[1])))
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, units, data = rows[0], rows[1], rows[2:]
data = data[stride - 1::stride]
ik = hdr.index('Kernel Name')
w = csv.writer(sys.stdout)
w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(data))])
w.writerow(['Kernel Name', ''] + [r[ik].replace('cfl::', '').replace('(bool)', '').replace('(int)', '')[:60] for r in data])
for m in WANT:
    if m in hdr:
        i = hdr.index(m)
        w.writerow([m, units[i]] + [r[i] for r in data])
