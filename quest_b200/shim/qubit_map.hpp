// qubit_map.hpp -- lazy qubit relabelling of GPU statevectors (quest_b200/shim/localiser_b200.cpp owns the state).
//
// The sharding logic keeps, per statevector Qureg (keyed by its device pointer), a permutation "logical qubit ->
// index bit".  Uncontrolled SWAPs only edit the permutation; a gate whose target sits on a rank bit pulls that
// qubit into the shard with ONE half-shard exchange and leaves it there (the reference swaps it in AND back:
// core/localiser.cpp:997-1040).  Everything that is not a relabelling-aware gate first restores the canonical
// order, so the permutation is never observable through QuEST's API.  These hooks let the other shim files
// restore / drop the permutation where amplitudes leave through a side door (gpu_copy*, gpu_sync, dealloc).
#ifndef QB_QUBIT_MAP_HPP
#define QB_QUBIT_MAP_HPP

void qbmap_canonicaliseHolding(const void* gpuPtr);   // the Qureg whose device amplitudes contain gpuPtr, if it is relabelled
void qbmap_forget(const void* gpuAmps);               // amplitudes are about to be freed or entirely overwritten
void qbmap_canonicaliseAll();                         // syncQuESTEnv(): user code may look at the device arrays next

#endif
