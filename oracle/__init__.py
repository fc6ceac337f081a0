"""CPU oracle for the STM E-step/M-step (TEST INFRASTRUCTURE — never imported by strutopy_b200).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
"""
