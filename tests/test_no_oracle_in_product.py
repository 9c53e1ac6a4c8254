"""CPU: the product package never imports, links or calls the oracle (or any CPU fallback)."""
import re
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
PRODUCT = REPO / "voxelized_geometry_tools_b200"


def test_product_sources_do_not_reference_the_oracle():
    offenders = []
    for path in list(PRODUCT.rglob("*.py")) + list(PRODUCT.rglob("*.cu")) + list(
            PRODUCT.rglob("*.cuh")) + list(PRODUCT.rglob("*.hpp")):
        text = path.read_text()
        if re.search(r"\boracle\b", text) or "libvgt_oracle" in text or "scipy" in text:
            offenders.append(str(path.relative_to(REPO)))
    assert offenders == []


def test_oracle_headers_say_test_infrastructure():
    for name in ("edt_oracle.cpp", "voxelizer_oracle.cpp", "oracle.py"):
        assert "TEST INFRASTRUCTURE ONLY" in (REPO / "oracle" / name).read_text()


def test_only_allowed_files_import_the_oracle():
    allowed = {"bench.py", "__graft_entry__.py"}
    for path in REPO.glob("*.py"):
        if "from oracle" in path.read_text() or "import oracle" in path.read_text():
            assert path.name in allowed, path.name
