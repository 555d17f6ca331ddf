/*
Copyright 2024 The RoleBasedGroup Authors.

Licensed under the Apache License, Version 2.0 (the "License").
*/

// Package placeroracle is the CPU ORACLE of the frozen placement spec (DESIGN.md §3) in Go, written
// in the reference's style so that a reviewer with a Go toolchain can run it next to the controller.
//
// TEST INFRASTRUCTURE — NOT sgl-project/rbg CODE and NOT COMPILED HERE (the build image has no Go
// toolchain; SURVEY.md §8c).  Upstream has no scoring / top-K / greedy path at all
// (pkg/scheduler only manages PodGroup CRs, pkg/scheduler/podgroup_manager.go:64-92): parity of
// placements is UNPINNED upstream.  This file is a line-for-line sibling of oracle/placer_oracle.c
// (the oracle the tests actually run) — the slow obvious restatement: dense role vectors, CSR rows
// walked in storage order with a float32 accumulator, a full sort for the top-K.
//
// What IS reference-defined and consumed here:
//   - replica order inside a step: roles in the given (lexicographic, pkg/dependency/dependency.go:133-137)
//     order, ordinals ascending (stateful_instance_set_utils.go:74-76);
//   - exclusive topology: all participating pods in ONE domain no other group occupies
//     (pkg/reconciler/pod_reconciler.go:172-231), per-role opt-out (annotation.go:29);
//   - gang = all-or-nothing (pkg/scheduler/k8s-scheduler-plugin/manager.go:131).
package placeroracle

import (
	"errors"
	"math"
	"sort"
)

const (
	FCap          = 8    // min(free, F)
	SelfW         = 8000 // self term = 8 x NVLink weight
	KMax          = 32
	MaxStepRoles  = 8
	MaxGroupRoles = 16

	StepExclusive = 1
	StepGang      = 2
	RoleExclusive = 1

	PlacedAll  = 0
	PlacedPart = 1
	GangFailed = 2
)

var (
	ErrInvalid = errors.New("placeroracle: malformed step")
	ErrInexact = errors.New("placeroracle: a score is not an exact integer below 2^24 (spec §3.4)")
)

// Topology is the cluster snapshot of spec §3.1: symmetric CSR with strictly ascending columns.
type Topology struct {
	RowPtr, ColIdx, EdgeW []int32
	Free, Domain          []int32
	DomainOwner           []int32 // -1 or the gid of the group that holds the domain exclusively
}

// Role is one role row of a step.
type Role struct {
	Count, Demand, Need int32
	Exclusive           bool
}

// Anchor is a pod of the group that is already placed: `Count` pods of group role Q on Node.
type Anchor struct{ Node, Q, Count int32 }

// Consumed is capacity the group already took in this reconcile and Free does not reflect yet.
type Consumed struct{ Node, Amount int32 }

// Step is one wave of one dependency level of one RoleBasedGroup (spec §3.2).
type Step struct {
	GID         int32
	Exclusive   bool
	Gang        bool
	FixedDomain int32 // -1 = none yet
	Roles       []Role
	Pair        [][]int32 // [len(Roles)][Q]
	Anchors     []Anchor
	Consumed    []Consumed
}

// Result of a step: Assign in replica order (-1 = unplaced), Status, exclusive Domain (-1 = none),
// and — for inspection — the dense (replica x node) Matrix and the per-role top-K key lists.
type Result struct {
	Assign []int32
	Status int32
	Domain int32
	Matrix [][]float32
	TopK   [][KMax]uint64
}

// orderableU32: monotone float32 -> uint32 map (spec §3.5).
func orderableU32(x float32) uint32 {
	b := math.Float32bits(x)
	if b>>31 != 0 {
		return b ^ 0xFFFFFFFF
	}
	return b ^ 0x80000000
}

// MakeKey: (score desc, node asc) as ONE descending uint64 key.
func MakeKey(s float32, node int32) uint64 {
	return uint64(orderableU32(s))<<32 | uint64(0xFFFFFFFF-uint32(node))
}

// KeyNode recovers the node of a key.
func KeyNode(k uint64) int32 { return int32(0xFFFFFFFF - uint32(k&0xFFFFFFFF)) }

// Place runs one step, literally per DESIGN.md §3.
func Place(t *Topology, st *Step) (*Result, error) {
	n := int32(len(t.RowPtr) - 1)
	P := len(st.Roles)
	if P < 1 || P > MaxStepRoles {
		return nil, ErrInvalid
	}
	Q := 0
	if len(st.Pair) > 0 {
		Q = len(st.Pair[0])
	}
	if Q > MaxGroupRoles {
		return nil, ErrInvalid
	}
	R := int32(0)
	for _, r := range st.Roles {
		if r.Count < 1 || r.Demand < 0 || r.Need < 0 {
			return nil, ErrInvalid
		}
		R += r.Count
	}
	if R > KMax {
		return nil, ErrInvalid
	}
	negInf := float32(math.Inf(-1))

	// dense anchor[q][m] and consumed[m] from the sparse records
	anchor := make([][]int32, Q)
	for q := range anchor {
		anchor[q] = make([]int32, n)
	}
	cons := make([]int32, n)
	for _, a := range st.Anchors {
		if a.Node < 0 || a.Node >= n || a.Q < 0 || int(a.Q) >= Q || a.Count < 0 {
			return nil, ErrInvalid
		}
		anchor[a.Q][a.Node] += a.Count
	}
	for _, c := range st.Consumed {
		if c.Node < 0 || c.Node >= n || c.Amount < 0 {
			return nil, ErrInvalid
		}
		cons[c.Node] += c.Amount
	}

	S := make([][]float32, P)
	A := make([]float32, n)
	for p, role := range st.Roles {
		roleExcl := st.Exclusive && role.Exclusive
		// role vector: A[m] = sum_q pair[p][q]*anchor[q][m] + need * min(free[m], F)
		for m := int32(0); m < n; m++ {
			v := int32(0)
			for q := 0; q < Q; q++ {
				v += st.Pair[p][q] * anchor[q][m]
			}
			f := t.Free[m]
			if f > FCap {
				f = FCap
			}
			A[m] = float32(v + role.Need*f)
		}
		// score: CSR row in storage order, float32 accumulator, then the self term
		S[p] = make([]float32, n)
		for i := int32(0); i < n; i++ {
			acc := float32(0)
			for j := t.RowPtr[i]; j < t.RowPtr[i+1]; j++ {
				acc += float32(t.EdgeW[j]) * A[t.ColIdx[j]]
			}
			acc += float32(SelfW) * A[i]
			if !(acc < 16777216.0) {
				return nil, ErrInexact
			}
			feasible := t.Free[i]-cons[i] >= role.Demand
			if roleExcl {
				o := t.DomainOwner[t.Domain[i]]
				feasible = feasible && (o == -1 || o == st.GID)
			}
			if feasible {
				S[p][i] = acc
			} else {
				S[p][i] = negInf
			}
		}
	}
	res := &Result{Domain: -1}
	for p, role := range st.Roles { // replicas of a role share the role row
		for c := int32(0); c < role.Count; c++ {
			res.Matrix = append(res.Matrix, S[p])
		}
	}

	// exclusive domain: fixed, or the domain of the best node of the FIRST participating role
	dstar := int32(-1)
	if st.Exclusive {
		if st.FixedDomain >= 0 {
			dstar = st.FixedDomain
		} else {
			for p, role := range st.Roles {
				if !role.Exclusive {
					continue
				}
				best, bn := uint64(0), int32(-1)
				for i := int32(0); i < n; i++ {
					if S[p][i] == negInf {
						continue
					}
					if k := MakeKey(S[p][i], i); k > best {
						best, bn = k, i
					}
				}
				if bn >= 0 {
					dstar = t.Domain[bn]
				}
				break
			}
		}
	}
	res.Domain = dstar

	// selection: top-K_p feasible nodes per role row, K_p = min(n, replicas up to and including role p)
	lists := make([][KMax]uint64, P)
	kacc := int32(0)
	for p, role := range st.Roles {
		kacc += role.Count
		K := kacc
		if K > n {
			K = n
		}
		roleExcl := st.Exclusive && role.Exclusive
		var keys []uint64
		for i := int32(0); i < n; i++ {
			if S[p][i] == negInf || (roleExcl && t.Domain[i] != dstar) {
				continue
			}
			keys = append(keys, MakeKey(S[p][i], i))
		}
		sort.Slice(keys, func(a, b int) bool { return keys[a] > keys[b] }) // keys are unique: any sort agrees
		for k := int32(0); k < K && int(k) < len(keys); k++ {
			lists[p][k] = keys[k]
		}
	}
	res.TopK = lists

	// greedy in replica order on a working copy of the capacity
	avail := make([]int32, n)
	for i := range avail {
		avail[i] = t.Free[i] - cons[i]
	}
	unplaced := 0
	for p, role := range st.Roles {
		for c := int32(0); c < role.Count; c++ {
			pick := int32(-1)
			for k := 0; k < KMax; k++ {
				key := lists[p][k]
				if key == 0 {
					break
				}
				if node := KeyNode(key); avail[node] >= role.Demand {
					pick = node
					break
				}
			}
			if pick >= 0 {
				avail[pick] -= role.Demand
			} else {
				unplaced++
			}
			res.Assign = append(res.Assign, pick)
		}
	}
	switch {
	case unplaced > 0 && st.Gang:
		for i := range res.Assign {
			res.Assign[i] = -1
		}
		res.Status = GangFailed
	case unplaced > 0:
		res.Status = PlacedPart
	default:
		res.Status = PlacedAll
	}
	return res, nil
}
