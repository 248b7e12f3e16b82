"""CPU oracle (test infrastructure only; never imported by gmvae_b200/)."""
