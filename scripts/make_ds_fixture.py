"""GPU box: `from_msa` on the reference's sample_example + two small fixtures -> the table-shaped update_DS archive
committed as tests/golden/update_DS_fixture.zip (CPU tests load it and export it to the reference's pickle layout)."""
import sys, shutil, tempfile
from argparse import Namespace
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from make_prg_b200.subcommands import from_msa
from make_prg_b200.subcommands.output_type import OutputType
tmp = Path(tempfile.mkdtemp())
src = tmp / "in"
src.mkdir()
ref = REPO / "tests" / "golden" / "ref"
for f in list((ref / "sample_example").glob("*.fa")) + [ref / "nested_snps_deletion.fa", ref / "contains_RYKMSW.fa"]:
    shutil.copy(f, src / f.name)
opts = Namespace(input=str(src), suffix="", output_prefix=str(tmp / "fx"), alignment_format="fasta", max_nesting=5,
                 min_match_length=7, output_type=OutputType("a"), force=True, threads=1, verbose=False, log=None, gpus=1)
from_msa.run(opts)
out = REPO / "gpurun_out" / "update_DS_fixture.zip"
shutil.copy(tmp / "fx.update_DS.zip", out)
shutil.copy(tmp / "fx.prg.fa", REPO / "gpurun_out" / "update_DS_fixture.prg.fa")
print("wrote", out, out.stat().st_size)
