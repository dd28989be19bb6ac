"""CPU oracle for the MelGAN hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package (see oracle/melgan_oracle.c for the parity-pinning statement).
"""
