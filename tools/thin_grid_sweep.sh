for W in 32768 49152 24576; do
echo "== W=$W default (auto)"; python tools/profile_walk.py 100000 $W 0 -1
echo "== W=$W static bs128"; python tools/profile_walk.py 100000 $W 128 0
echo "== W=$W static bs64"; python tools/profile_walk.py 100000 $W 64 0
echo "== W=$W static bs32"; python tools/profile_walk.py 100000 $W 32 0
nb=$((W/128))
echo "== W=$W dyn grid=nb128($nb)"; MCIG_DYN_GRID=$nb python tools/profile_walk.py 100000 $W 0 1
echo "== W=$W dyn bs64 (auto grid)"; MCIG_DYN_BS=64 python tools/profile_walk.py 100000 $W 0 1
echo "== W=$W dyn bs32 (auto grid)"; MCIG_DYN_BS=32 python tools/profile_walk.py 100000 $W 0 1
echo "== W=$W dyn bs32 grid=all"; MCIG_DYN_BS=32 MCIG_DYN_GRID=$((W/32)) python tools/profile_walk.py 100000 $W 0 1
done
