#!/bin/bash
# r01z (1 GPU): final check at HEAD -- GPU suite, cfg-5 full-size test + probe, bench both arms
TAG=${1:-r01z}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_relocalisation.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 150 python -m pytest tests/test_gpu_relocalisation.py -x -q --durations=3 > gpurun_out/${TAG}_pytest_reloc.log 2>&1; echo "reloc rc=$?"; tail -8 gpurun_out/${TAG}_pytest_reloc.log
timeout 150 python scripts/reloc_probe.py > gpurun_out/${TAG}_reloc_probe.json 2> gpurun_out/${TAG}_reloc_probe.err; echo "probe rc=$?"; cat gpurun_out/${TAG}_reloc_probe.json; tail -3 gpurun_out/${TAG}_reloc_probe.err
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
