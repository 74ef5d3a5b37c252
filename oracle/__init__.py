"""TEST INFRASTRUCTURE — CPU oracle of the Latent-Paint render path (see kaolin_shim.py).
Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
