echo "90 runs prefetch on"; timeout 200 python tools/probe_launch.py 90 1
echo "90 runs prefetch off"; timeout 200 python tools/probe_launch.py 90 0
echo "30 runs"; timeout 200 python tools/probe_launch.py 30 1
echo "10 runs"; timeout 200 python tools/probe_launch.py 10 1
