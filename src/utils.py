"""reference: src/utils.py -- the functions gnomix.py's inference path imports."""
from gnomix_b200.io import read_vcf, vcf_to_npy, snp_intersection, read_genetic_map  # noqa: F401
from gnomix_b200.cli import npy_to_vcf, update_vcf, read_headers  # noqa: F401
