"""Writes the inputs of the JVM golden run (tools/make_jvm_golden.sh): seeded DBs as .rgdb + reads as FASTA.
No GPU needed.  usage: python make_inputs.py OUT_DIR"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rappas_b200 import _abi, synth  # noqa: E402
from rappas_b200._lib import check, load  # noqa: E402

CASES = {
    "nucl_k8": (dict(alphabet=0, k=8, n_nodes=299, n_keys=49152, mean_postings=16, seed=43),
                dict(n_reads=500, length=(20, 300), seed=1043, iupac_rate=0.004, n_rate=0.002, gap_rate=0.001)),
    "nucl_k10_ties": (dict(alphabet=0, k=10, n_nodes=1999, n_keys=100000, mean_postings=32, seed=44),
                      dict(n_reads=500, length=150, seed=1044)),
    "amino_k4": (dict(alphabet=1, k=4, n_nodes=77, n_keys=5000, mean_postings=7, seed=11),
                 dict(n_reads=300, length=40, seed=12, mutation=0.05, iupac_rate=0.02)),
}


def main(out):
    fn = load()
    for name, (dbkw, rkw) in CASES.items():
        db = synth.make_db(**dbkw)
        rb = synth.make_reads(db, **rkw)
        desc = _abi.RpDbDesc(db.alphabet, db.k, db.n_nodes, float(db.thr_log10), float(db.thr_lin), 0, db.n_keys, db.n_postings)
        check(fn["db_save_file"](os.path.join(out, name + ".rgdb").encode(), C.byref(desc), _abi.ptr(db.keys),
                                 _abi.ptr(db.offsets), _abi.ptr(db.post_node), _abi.ptr(db.post_score)))
        with open(os.path.join(out, name + ".fasta"), "w") as f:
            for i in range(rb.n_reads):
                f.write(">r%d\n%s\n" % (i, rb.read(i)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
