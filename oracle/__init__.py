"""CPU oracle (test infrastructure only): see frontend_oracle.c.  Imported only by tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke()."""
