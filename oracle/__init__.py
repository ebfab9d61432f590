"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's naive backends plus (oracle/_ref) the reference's own generated
naive code compiled from /root/reference.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product (dawn_b200/) never does.
"""
