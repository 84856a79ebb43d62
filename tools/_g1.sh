mkdir -p gpurun_out
NTB_TUNE_E2E=1 python tools/scan_tune.py '' 'NTB_CONTIG_GROUP_RATIO=0.12' 'NTB_CONTIG_GROUP_RATIO=0.07' 'NTB_CONTIG_GROUPS=3,NTB_CONTIG_GROUP_RATIO=0.28' > gpurun_out/r02ai_ratio.log 2> gpurun_out/r02ai_ratio.err
cat gpurun_out/r02ai_ratio.log
