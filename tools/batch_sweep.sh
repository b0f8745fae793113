for b in 131072 262144 524288 1048576 2097152 4194304 8388608; do echo "batch $b"; python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 2 --batch $b | cut -c40-; done
