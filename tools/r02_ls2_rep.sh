for i in 1 2 3; do for nw in 1 2 4; do echo -n "NW=$nw "; FM_SCAN_BWD_LS2_NW=$nw timeout 120 tools/ubench/ls2_check 8 4096 | tail -1; done; done
