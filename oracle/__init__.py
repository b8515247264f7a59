"""CPU oracle of the roadmap-construction path.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product
package.  See oracle/grr_oracle.c for the parity status ("parity unpinned" at the Klamp't/IKDB
boundary; workspace graphs pinned by the reference's shipped fixture)."""
