"""CPU oracle for the PdsNetwork.forward hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product package
(practicaldeepstereo_nips2018_b200) never does.
"""
