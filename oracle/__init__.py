"""CPU oracle for the duwu diffusion training step — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker.  The product (uwudiff_b200/) never imports it and has no CPU fallback.

Parity status: the loss/noising path (oracle/loss_oracle.py) is pinned against the reference's own
`src/duwu/loss/diffusion.py` executed verbatim in the build container (oracle/ref_loss.py +
oracle/make_golden.py -> tests/golden/loss_*.npz).  The denoiser (oracle/unet_oracle.py) and LyCORIS
(oracle/lycoris_oracle.py) restate un-vendored third-party packages (diffusers, lycoris-lora) that are absent
from /root/reference and from this image: for those, PARITY IS UNPINNED.
"""
