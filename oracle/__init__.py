"""CPU oracle for the PTT point-feature hot path.  TEST INFRASTRUCTURE -- not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under ptt_b200/ does.  See oracle/pointnet2_ref.c for the parity
status of the native ops ("parity unpinned") and oracle/torch_port.py for the Python half.
"""
