for plan in 1:32 2:32 4:32 1:64 2:64 4:64 1:96 2:96 1:128 2:128 4:128; do echo -n "plan=$plan  "; VBX_LPC_PREFETCH=0 VBX_LPC_PLAN=$plan python tools/lpc_split.py; done
