for sh in stage1 stage2 stage3; do
  echo "== $sh default"; python tools/tune.py --shape $sh --dtype f32 --fwd "" --fwd16 "0x0x0" --bwd "0x0x1" 2>&1 | grep -E "us|error"
done
echo "== stage1 sweeps"; python tools/tune.py --shape stage1 --dtype f32 --fwd "" --fwd16 "2x4x1,4x4x1,4x2x1,4x4x2" --bwd "32x8x1,16x4x1,8x4x1,4x4x1,8x2x1" 2>&1 | grep -E "us|error"
echo "== stage2 sweeps"; python tools/tune.py --shape stage2 --dtype f32 --fwd "" --fwd16 "2x4x1,4x4x1,4x2x1,4x1x1" --bwd "8x4x1,4x4x1,8x2x1,4x2x1,2x1x1" 2>&1 | grep -E "us|error"
echo "== stage3 sweeps"; python tools/tune.py --shape stage3 --dtype f32 --fwd "" --fwd16 "2x4x1,4x4x1,4x2x1,4x1x1" --bwd "2x1x1,4x1x1,4x2x1" 2>&1 | grep -E "us|error"
