"""Region builder on chr20-scale call sets: device (through the C ABI, host call sets -> resident batch) vs
the C++ restatement (1 thread, like the reference's single-threaded builder, main.rs:216) vs the generator's Python builder."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import oracle_py as orc
from aardvark_b200 import synth
from aardvark_b200.batch import CallSets, CompareOutputs
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
L = int(synth.CHR20_LEN * scale)
ref, inputs = synth.callsets_compare(L, synth.SynthParams(n_variants=int(150_000 * scale)), 20)
cs = CallSets(inputs)
s = Solver(0); s.set_reference([ref])
for _ in range(3):
    t0 = time.perf_counter(); n, nv = s.build_regions(cs, 0, 50, download=False); dt_dev = time.perf_counter() - t0
t0 = time.perf_counter(); o = orc.build_regions(cs, L, 50); dt_orc = time.perf_counter() - t0
t0 = time.perf_counter(); h = synth.cluster_regions(inputs, L, 50); dt_py = time.perf_counter() - t0
d = s.build_regions(cs, 0, 50)
same = all(np.array_equal(getattr(d, f), getattr(o, f)) for f in ("region_id", "start", "end", "var_off", "position", "zygosity", "allele_off", "a0_len", "a1_len"))
print(f"{cs.n_variants} variants in {cs.n_inputs} call sets -> {n} clusters; device builder {dt_dev * 1e3:.2f} ms (incl. H2D of the call sets), "
      f"C++ restatement {dt_orc * 1e3:.1f} ms, Python host builder {dt_py * 1e3:.0f} ms; device == restatement: {same}")
s.build_regions(cs, 0, 50, download=False)
s.run_resident(CompareConfig(enable_sequences=False))
a = s.download(CompareOutputs(h, region_metrics=False))
b = s.compare_batch(h, CompareConfig(enable_sequences=False), region_metrics=False)
print("solve on the device-built batch == solve on the uploaded host-built batch:", a.diff(b) == [])
