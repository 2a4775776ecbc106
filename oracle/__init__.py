"""CPU oracle for the RetinaNet loss / post-processing path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product package (neuralnetworklibrary_b200) never does.
"""
