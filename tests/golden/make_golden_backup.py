#!/usr/bin/env python
"""Generates tests/golden/backup/*.json: the texts the REFERENCE's own BackupRestoreFacility<YaspGrid>::backup writes
(dune/grid/yaspgrid/backuprestore.hh compiled in place into oracle/_ref), one file per case of
tests/test_backup_restore.py::CASES with the configuration and the text of every rank. Run in the build container:

    make -C oracle ref && python tests/golden/make_golden_backup.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402
from tests.common import oracle_world  # noqa: E402
from tests.test_backup_restore import CASES  # noqa: E402


def main():
    o = orc.load("ref")
    assert o.kind == "reference"
    os.makedirs(os.path.join(HERE, "backup"), exist_ok=True)
    for k, (cfg, P) in enumerate(CASES):
        w = oracle_world(o, cfg, P)
        texts = [w.backup(r) for r in range(P)]
        w.close()
        c = dict(cfg)
        if c["tensor"] is not None:
            c["tensor"] = [[float(x) for x in t] for t in c["tensor"]]
        path = os.path.join(HERE, "backup", f"case{k:02d}.json")
        json.dump({"cfg": c, "P": P, "texts": texts}, open(path, "w"), indent=1)
        print(path, sum(len(t) for t in texts), "bytes of text")


if __name__ == "__main__":
    main()
