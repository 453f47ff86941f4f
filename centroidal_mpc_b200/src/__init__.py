"""Host-side mirror of the reference's ``src`` package for the SCP hot path
(same module and function names as /root/reference/src)."""
