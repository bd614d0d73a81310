"""Summarise an ncu --set full report: headline metrics + stall samples by kernel role (for the warp-specialised
pointwise-conv kernels the roles are separated by landmarks in the SASS stream).

    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [--top 20]
"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__cycles_elapsed.avg',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic']
STALLS = ['stall_long_sb', 'stall_wait', 'stall_selected', 'stall_short_sb', 'stall_mio', 'stall_lg', 'stall_barrier',
          'stall_branch_resolving', 'stall_math', 'stall_not_selected', 'stall_no_inst', 'stall_dispatch', 'stall_sleep',
          'stall_membar', 'stall_tex', 'stall_drain', 'stall_misc']


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 15
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("   %-80s %s %s" % (w, r[i], units[i]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    si, wi, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    idx = [hdr.index(n) for n in STALLS]
    data = [(int(r[wi] or 0), r[si].strip(), int(r[ii] or 0), [int(r[i] or 0) for i in idx]) for r in rows[2:] if len(r) > wi]
    tot = sum(d[0] for d in data) or 1
    toti = sum(d[2] for d in data) or 1
    # role boundaries: first BAR.SYNC (end of setup), first LDTM-ish / UTCBAR etc.
    def first(pred, start=0):
        for n in range(start, len(data)):
            if pred(data[n][1]):
                return n
        return len(data)
    b0 = first(lambda s: 'BAR.SYNC' in s)
    mma_end = first(lambda s: 'UTCBAR' in s, b0)
    mma_end = max([n for n, d in enumerate(data) if 'UTCBAR' in d[1]] or [b0])
    ldtm0 = first(lambda s: 'LDTM' in s, b0)
    b1 = first(lambda s: 'BAR.SYNC' in s, b0 + 1)
    # epilogue ends at the tmem_empty arrive after the last LDTM; approximate with the first SYNCS.ARRIVE after the last STG
    last_ldtm = max([n for n, d in enumerate(data) if 'LDTM' in d[1]] or [b0])
    epi_end = first(lambda s: 'SYNCS.ARRIVE' in s, last_ldtm)
    regions = {"setup": (0, b0 + 1), "mma": (b0 + 1, mma_end + 1), "epilogue": (mma_end + 1, epi_end + 1),
               "producer": (epi_end + 1, b1), "teardown+wait loops": (b1, len(data))}
    print("stall samples %d, warp instructions %d" % (tot, toti))
    for name, (a, b) in regions.items():
        s = sum(d[0] for d in data[a:b])
        ins = sum(d[2] for d in data[a:b])
        st = [sum(d[3][k] for d in data[a:b]) for k in range(len(STALLS))]
        tops = sorted(zip(STALLS, st), key=lambda t: -t[1])[:4]
        print("  %-20s sass[%5d,%5d) samples %6d (%4.1f%%) warp-instr %11d (%4.1f%%) %s" % (name, a, b, s, 100 * s / tot, ins, 100 * ins / toti, tops))
    print("top stalled instructions:")
    for n, d in sorted(enumerate(data), key=lambda t: -t[1][0])[:top]:
        st = sorted(zip(STALLS, d[3]), key=lambda t: -t[1])[:2]
        print("  %5d %6d %4.1f%% exec %9d %-66s %s" % (n, d[0], 100 * d[0] / tot, d[2], d[1][:66], st))


if __name__ == "__main__":
    main()
