"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with gpurun) so that
`bench.py --impl reference` and tests/test_dropin_pipeline.py can run its own classes there.  Build container only.

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>

The install is made from a copy under /tmp because /root/reference is read-only and setuptools_scm writes _version.py into
the tree.  One packaging fix is applied to the COPY's pyproject.toml: `include = ["mosaicking"]` becomes
`["mosaicking*"]` -- as shipped the wheel omits the `mosaicking.core` subpackage and does not import.  No source file is
touched.  --no-deps: `shapely` is not in the offline wheelhouse (oracle/ref_shim.py stubs the one call that needs it).
The reference's small test assets (tests/data: one 830 KB video + its orientation log) are copied next to it for the
drop-in pipeline test."""
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def main(force: bool = False) -> bool:
    dst = ROOT / "baseline" / "_ref"
    if not (REF / "pyproject.toml").exists():
        return (dst / "mosaicking" / "mosaic.py").exists()
    if (dst / "mosaicking" / "core" / "interface.py").exists() and (dst / "testdata" / "sparus_snippet.mp4").exists() and not force:
        return True
    tmp = Path("/tmp/mosaic_reference_copy")
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns(".git"))
    pp = tmp / "pyproject.toml"
    pp.write_text(pp.read_text().replace('include = ["mosaicking"]', 'include = ["mosaicking*"]'))
    shutil.rmtree(dst, ignore_errors=True)
    dst.parent.mkdir(parents=True, exist_ok=True)
    subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                           "--find-links", "/opt/wheelhouse", "--target", str(dst), str(tmp)])
    (dst / "testdata").mkdir(exist_ok=True)
    for name in ("sparus_snippet.mp4", "sparus_snippet.csv", "sparus_time_offset.yaml"):
        shutil.copy(REF / "tests" / "data" / name, dst / "testdata" / name)
    return True


if __name__ == "__main__":
    print("baseline/_ref installed:", main(force="--force" in sys.argv))
