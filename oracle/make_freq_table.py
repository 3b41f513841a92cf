"""Extracts the reference's DATA_FREQ_TABLE (datasets/encode_openx_dataset.py:51-108: dataset name -> control frequency in
Hz, read by hma/data.py:207 to derive the frame stride) into hma_b200/data_freq_table.json without importing the module
(it pulls in tensorflow_datasets). Run in the authoring container: python oracle/make_freq_table.py"""
import ast
import json
from pathlib import Path

SRC = Path("/root/reference/datasets/encode_openx_dataset.py")
DST = Path(__file__).resolve().parent.parent / "hma_b200" / "data_freq_table.json"

tree = ast.parse(SRC.read_text())
table = None
for node in tree.body:
    if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "DATA_FREQ_TABLE" for t in node.targets):
        table = ast.literal_eval(node.value)
assert isinstance(table, dict) and len(table) > 40
DST.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
print(f"{len(table)} entries -> {DST}")
