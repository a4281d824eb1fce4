for b in 4 6 7 8 12 16 24; do timeout 120 tools/ubench/ls2_check $b 4096 | tail -1; done
