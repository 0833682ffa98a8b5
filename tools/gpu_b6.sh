#!/bin/bash
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_bytes.sum
for k in 0 3; do
STEIN_PANEL_DEBUG_SKIP=$k STEIN_SKIP_MEDIAN=1 ncu --metrics $M --clock-control none -k regex:'panel_gemm_kernel' -s 1 -c 2 --csv --log-file gpurun_out/panel_skip$k.csv \
    python tools/panel_bench.py 32768 1024 1 > gpurun_out/ncu_panel_$k.log 2>&1; echo "ncu_rc=$?"
done
