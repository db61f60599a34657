"""physdock_b200 -- B200-native (sm_100a) implementation of PhysDock's reverse-diffusion sampling step."""
