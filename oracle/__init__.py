"""CPU oracle of the embedding-distance path.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under deep_insight_face_b200/ imports this package (tests/test_abi.py::test_product_never_imports_the_oracle checks that).
"""
