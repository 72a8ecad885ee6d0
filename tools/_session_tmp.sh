mkdir -p gpurun_out/r2b6
timeout 120 python tools/chain_time.py 4 > gpurun_out/r2b6/time_default.txt 2>&1; sed -n 5,8p gpurun_out/r2b6/time_default.txt
MODA_B200_LIB=gpurun_variants/lib_aliaspe.so timeout 120 python tools/chain_time.py 4 > gpurun_out/r2b6/time_aliaspe.txt 2>&1; sed -n 5,8p gpurun_out/r2b6/time_aliaspe.txt
