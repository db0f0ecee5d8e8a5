"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's Chamfer/DCD algorithm, used as the parity checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The shipped package must never import anything from here.
"""
