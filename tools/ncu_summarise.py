#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` of one step's map-pipeline launches into profiles/map_kernel_traffic.json.

    python tools/ncu_summarise.py gpurun_out/r01y_full_raw.csv profiles/map_kernel_traffic.json "source text" [maps]
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
	src, dst, source = sys.argv[1], sys.argv[2], sys.argv[3]
	maps = int(sys.argv[4]) if len(sys.argv) > 4 else 20743
	rows = list(csv.reader(open(src)))
	hdr, units = rows[0], rows[1]

	def get(r, name, default=None):
		if name not in hdr:
			return default
		v = r[hdr.index(name)].replace(',', '')
		try:
			return float(v)
		except ValueError:
			return v

	def mbytes(r, name):
		v = get(r, name, 0.0)
		u = units[hdr.index(name)]
		return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[u]

	def msec(r, name):
		v = get(r, name, 0.0)
		u = units[hdr.index(name)]
		return v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1.0)
	ks = []
	rd = wr = 0.0
	for r in rows[2:]:
		k = dict(kernel=get(r, 'Kernel Name'), grid=get(r, 'Grid Size'), block=get(r, 'Block Size'),
				ms=msec(r, 'gpu__time_duration.sum'),
				dram_read_MB=mbytes(r, 'dram__bytes_read.sum'), dram_write_MB=mbytes(r, 'dram__bytes_write.sum'),
				ipc=get(r, 'sm__inst_executed.avg.per_cycle_elapsed'),
				issue_active_pct=get(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active'),
				warps_active_pct=get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
				dram_pct=get(r, 'dram__bytes.sum.pct_of_peak_sustained_elapsed', get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')),
				regs=get(r, 'launch__registers_per_thread'), inst=get(r, 'sm__inst_executed.sum'))
		rd += k['dram_read_MB'] * 1e6
		wr += k['dram_write_MB'] * 1e6
		ks.append(k)
	import bench
	out = dict(source=source, kernel_source_hash=bench._kernel_source_hash(), maps=maps, dram_bytes_per_step=rd + wr, dram_bytes_read=rd, dram_bytes_written=wr,
			dram_bytes_per_map=(rd + wr) / maps, serialised_ms=sum(k['ms'] for k in ks), kernels=ks)
	json.dump(out, open(dst, 'w'), indent=1)
	print('%d launches, %.1f ms serialised, %.1f MB read, %.1f MB written, %.0f B/map' % (len(ks), out['serialised_ms'], rd / 1e6, wr / 1e6, out['dram_bytes_per_map']))


if __name__ == '__main__':
	main()
