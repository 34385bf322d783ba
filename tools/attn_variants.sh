#!/bin/bash
# A/B builds of the streaming attention kernel's tunables (run HERE: nvcc cross-compiles), each into its own .so;
# on the GPU box: SHOWTELL_B200_LIBNAME=libshowtell_b200_vN.so python bench.py --workload attn_gru_train ...
set -e
build() {  # name flags
  SHOWTELL_B200_LIBNAME=libshowtell_b200_$1.so SHOWTELL_B200_VARIANT="attn_stream.cu:$2" python -m showtell_b200.build --relink | tail -1
}
build v1 "-DST_ATTN_NCW=16"
build v2 "-DST_ATTN_NCW=8 -DST_ATTN_STG=5 -DST_ATTN_CTAS_PER_SM=2"
build v3 "-DST_ATTN_NCW=16 -DST_ATTN_STG=6 -DST_ATTN_CHUNK=32768"
build v4 "-DST_ATTN_NCW=8 -DST_ATTN_STG=24 -DST_ATTN_CHUNK=8192"
build v5 "-DST_ATTN_NCW=12 -DST_ATTN_STG=4 -DST_ATTN_CTAS_PER_SM=2 -DST_ATTN_CHUNK=16384"
