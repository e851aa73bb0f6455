"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the MM-RCA late-fusion head of
espiriki/Garbage_Classification_RCA (CVPR_code/multimodal_model.py:636-728).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product package
(garbage_classification_rca_b200) never does.
"""
