//go:build cgo && placement_gpu

// scheduler_gpu.go -- the cgo shim that puts the B200 placement engine (libplacement.so, include/placement_engine.h)
// behind SwarmKit's manager/scheduler.  Drop this file next to manager/scheduler/scheduler.go and build the manager
// with `-tags placement_gpu`; scheduler.go changes in three places only (INTEGRATION.md lists them):
//
//   New():                   s.gpu = mustNewGPUEngine()
//   tick():                  the two scheduleTaskGroup loops (scheduler.go:464-469) -> s.scheduleTickGPU(ctx, groups, decisions)
//   processPreassignedTasks: taskFitNode (scheduler.go:646-690)                  -> s.taskFitNodeGPU(ctx, t, nodeID)
//
// Everything else -- the store, the watch loop, applySchedulingDecisions, NodeInfo bookkeeping -- is the reference's own
// code.  The reference scheduler is one goroutine (scheduler.go:175-237), and so is the engine's contract: no locking here.
//
// This file is the Go statement of swarmkit_b200/csrc/scheduler_host.cpp, which is what this repository compiles and
// tests (there is no Go toolchain in its build image): same dictionaries, same encoders, same call sequence.  Function
// by function:   encodeRow <- Scheduler::encode_row   encodeGroup <- encode_group   compileConstraints <-
// compile_constraints   scheduleTickGPU <- scheduleTaskGroups/scheduleRun   schedulePreferenceGroup/fillLeaf/
// scheduleNTasksOnSubtreeGPU <- the functions of the same names   taskFitNodeGPU <- processPreassignedTasks / fit_run
// scheduleVolumeGroup / scheduleVolumeGroupStepwise / fitMany / countExcluded / volumesStaticFor <- scheduleVolumeGroup /
// scheduleVolumeGroupStepwise / fit_many / count_excluded / VolumeBook::staticFor
// (the volume bookkeeping itself is the reference's own volumeSet, volumes.go).
package scheduler

/*
#cgo LDFLAGS: -lplacement
#include <stdlib.h>
#include "placement_engine.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"net"
	"sort"
	"strings"
	"time"
	"unsafe"

	"github.com/moby/swarmkit/v2/api"
	"github.com/moby/swarmkit/v2/api/genericresource"
	"github.com/moby/swarmkit/v2/log"
	"github.com/moby/swarmkit/v2/manager/constraint"
	"github.com/moby/swarmkit/v2/manager/state/store"
	"github.com/moby/swarmkit/v2/protobuf/ptypes"
)

// gpuEngine owns the device mirror of nodeSet and the string dictionaries (SURVEY Appendix B): everything that
// crosses the C ABI is an integer.
type gpuEngine struct {
	h         *C.pe_engine
	values    map[string]uint32 // folded constraint operands / attribute values; "" = 0 (strings.EqualFold -> ==)
	exact     map[string]uint32 // Platform.OS / normalised Platform.Architecture (filter.go:291-312)
	services  map[string]uint32      // service ID -> id of its dense counter column in the engine (ids are recycled, see serviceID)
	freeSvc   []uint32
	nextSvc   uint32
	recycleAt int
	kinds     map[string]uint32      // generic resource kinds
	labelCols map[string]uint32      // "n:<key>" / "e:<key>" -> attribute column of FOLDED values (constraints)
	prefCols  map[string]uint32      // same keys -> attribute column of EXACT values (placement preferences)
	prefIDs   map[string]uint32      // exact label value -> id, "" = 0
	prefStr   []string               // id -> value (canonical branch order needs the strings back)
	ports     map[hostPortSpec]uint32 // (protocol, port) -> bit slot
	plugins   map[[2]string]uint32   // (type, name) -> bit slot
	nextCol   uint32
	order     []string            // row index -> node ID, ascending (the canonical tie-break, SURVEY 8c)
	index     map[string]uint32   // node ID -> row index
	dirty     map[string]struct{} // rows whose NodeInfo changed since the last upload
	layout    bool                // membership or a dictionary grew: every row is uploaded again
	// CSI cluster volumes: the node set VolumesFilter allows for the group being scheduled, as one attribute column
	volCol    uint32            // C.PE_NONE until the first group with cluster mounts
	volIn     map[string]struct{} // the nodes that carry volMark in that column (every other row holds 0 = "")
	leafAlsoInVolumeSet bool   // a preference group with cluster mounts: every leaf visit carries the volume term too
	leafStepwise        bool   // ... and, when its volumes count their users, walks the fill loop itself
}

func mustNewGPUEngine() *gpuEngine {
	cfg := C.pe_config{abi_version: C.PE_ABI_VERSION, device: -1, world_size: 1}
	var h *C.pe_engine
	if rc := C.pe_create(&cfg, &h); rc != C.PE_OK {
		// no CPU fallback: a manager built with this tag needs its GPU
		panic("placement engine: " + C.GoString(C.pe_last_error(nil)))
	}
	return &gpuEngine{h: h, values: map[string]uint32{"": 0}, exact: map[string]uint32{"": 0}, services: map[string]uint32{},
		kinds: map[string]uint32{}, labelCols: map[string]uint32{}, prefCols: map[string]uint32{}, prefIDs: map[string]uint32{"": 0},
		prefStr: []string{""}, ports: map[hostPortSpec]uint32{}, plugins: map[[2]string]uint32{}, nextCol: C.PE_ATTR_FIRST_LABEL,
		index: map[string]uint32{}, dirty: map[string]struct{}{}, layout: true, recycleAt: 4096, volCol: C.PE_NONE,
		volIn: map[string]struct{}{}}
}

// serviceID names the engine's per-service counter column.  The engine never frees a column, so the ids of services that
// no node counts and no task references any more are handed out again (recycleServiceIDs, called at the top of a tick):
// the columns are bounded by the services alive at one time.  A recycled column is all zeros: the device counts mirror
// NodeInfo.ActiveTasksCountByService.
func (e *gpuEngine) serviceID(s string) uint32 {
	if id, ok := e.services[s]; ok {
		return id
	}
	var id uint32
	if n := len(e.freeSvc); n > 0 {
		id, e.freeSvc = e.freeSvc[n-1], e.freeSvc[:n-1]
	} else {
		id = e.nextSvc
		e.nextSvc++
	}
	e.services[s] = id
	return id
}

func (s *Scheduler) recycleServiceIDs() {
	e := s.gpu
	if len(e.services) < e.recycleAt {
		return
	}
	live := map[string]struct{}{}
	for _, ni := range s.nodeSet.nodes {
		for svc, c := range ni.ActiveTasksCountByService {
			if c != 0 {
				live[svc] = struct{}{}
			}
		}
	}
	for _, t := range s.allTasks {
		live[t.ServiceID] = struct{}{}
	}
	for svc, id := range e.services {
		if _, ok := live[svc]; !ok {
			e.freeSvc = append(e.freeSvc, id)
			delete(e.services, svc)
		}
	}
	if e.recycleAt = 2 * len(e.services); e.recycleAt < 4096 {
		e.recycleAt = 4096
	}
}

func (e *gpuEngine) err(what string) error { return fmt.Errorf("%s: %s", what, C.GoString(C.pe_last_error(e.h))) }

func intern(m map[string]uint32, s string) uint32 {
	if id, ok := m[s]; ok {
		return id
	}
	id := uint32(len(m))
	m[s] = id
	return id
}

// fold is strings.EqualFold made into equality: pe_fold_value (A-Z, U+212A, U+017F) on both operands.
func (e *gpuEngine) fold(s string) string {
	if s == "" {
		return s
	}
	in := C.CString(s)
	defer C.free(unsafe.Pointer(in))
	out := make([]byte, len(s)+1)
	n := C.pe_fold_value(in, C.uint32_t(len(s)), (*C.char)(unsafe.Pointer(&out[0])), C.uint32_t(len(out)))
	return string(out[:n])
}
func (e *gpuEngine) valueID(s string) uint32 { return intern(e.values, e.fold(s)) }
func (e *gpuEngine) column(m map[string]uint32, key string) uint32 {
	if c, ok := m[key]; ok {
		return c
	}
	m[key] = e.nextCol
	e.nextCol++
	e.layout = true // a new column has to be filled for every node
	return m[key]
}
func (e *gpuEngine) prefID(v string) uint32 {
	if id, ok := e.prefIDs[v]; ok {
		return id
	}
	e.prefIDs[v] = uint32(len(e.prefStr))
	e.prefStr = append(e.prefStr, v)
	return e.prefIDs[v]
}
func (e *gpuEngine) pluginSlot(typ, name string) uint32 {
	k := [2]string{typ, name}
	if s, ok := e.plugins[k]; ok {
		return s
	}
	e.plugins[k] = uint32(len(e.plugins))
	e.layout = true
	return e.plugins[k]
}
func (e *gpuEngine) portSlot(p hostPortSpec) uint32 {
	if s, ok := e.ports[p]; ok {
		return s
	}
	e.ports[p] = uint32(len(e.ports))
	return e.ports[p]
}
func normArch(a string) string { // filter.go:291-306
	switch a {
	case "x86_64":
		return "amd64"
	case "aarch64":
		return "arm64"
	}
	return a
}

// ---- node rows: the device mirror of NodeInfo (nodeinfo.go:28-44) --------------------------------------------------

const volMark = 1 // the value of gpuEngine.volCol on the rows of the current volume node set

type rowBatch struct {
	rows  []C.pe_node_row
	attrs []C.pe_kv32
	svcs  []C.pe_kv32
	gens  []C.pe_kv64
	ports []C.uint32_t
	plugs []C.uint32_t
}

func (e *gpuEngine) encodeRow(idx uint32, ni *NodeInfo, b *rowBatch) {
	n := ni.Node
	r := C.pe_node_row{node_idx: C.uint32_t(idx), flags: C.PE_NODE_VALID}
	if n.Status.State == api.NodeStatus_READY && n.Spec.Availability == api.NodeAvailabilityActive { // filter.go:40-43
		r.flags |= C.PE_NODE_READY
	}
	d := n.Description
	if d != nil && d.Platform != nil {
		r.flags |= C.PE_NODE_HAS_PLATFORM
		r.os_id = C.uint32_t(intern(e.exact, d.Platform.OS))
		r.arch_id = C.uint32_t(intern(e.exact, normArch(d.Platform.Architecture)))
	}
	if d != nil && d.Engine != nil {
		r.flags |= C.PE_NODE_HAS_ENGINE
	}
	if ip := net.ParseIP(n.Status.Addr); ip != nil { // constraint.go:127-146
		r.flags |= C.PE_NODE_IP_VALID
		if ip.To4() != nil {
			r.flags |= C.PE_NODE_IP_V4
		}
		ip16 := ip.To16()
		for w := 0; w < 4; w++ {
			r.ip[w] = C.uint32_t(uint32(ip16[4*w])<<24 | uint32(ip16[4*w+1])<<16 | uint32(ip16[4*w+2])<<8 | uint32(ip16[4*w+3]))
		}
	}
	if ni.AvailableResources != nil {
		r.cpu_avail, r.mem_avail = C.int64_t(ni.AvailableResources.NanoCPUs), C.int64_t(ni.AvailableResources.MemoryBytes)
	}
	r.total_tasks = C.uint32_t(ni.ActiveTasksCount)
	r.attr_off = C.uint32_t(len(b.attrs))
	attr := func(col uint32, v uint32) {
		if v != 0 {
			b.attrs = append(b.attrs, C.pe_kv32{key: C.uint32_t(col), value: C.uint32_t(v)})
		}
	}
	hostname, os, arch := "", "", ""
	if d != nil {
		hostname = d.Hostname
		if d.Platform != nil {
			os, arch = d.Platform.OS, d.Platform.Architecture
		}
	}
	attr(C.PE_ATTR_NODE_ID, e.valueID(n.ID))       // constraint.go:110
	attr(C.PE_ATTR_HOSTNAME, e.valueID(hostname))  // :114-125
	attr(C.PE_ATTR_ROLE, e.valueID(n.Role.String())) // :147-150 (observed role)
	attr(C.PE_ATTR_OS, e.valueID(os))              // :151-160
	attr(C.PE_ATTR_ARCH, e.valueID(arch))          // :161-170 (NOT normalised)
	label := func(key string) (string, bool) {
		if key[0] == 'n' {
			v, ok := n.Spec.Annotations.Labels[key[2:]]
			return v, ok
		}
		if d != nil && d.Engine != nil {
			v, ok := d.Engine.Labels[key[2:]]
			return v, ok
		}
		return "", false
	}
	for key, col := range e.labelCols {
		if v, ok := label(key); ok {
			attr(col, e.valueID(v))
		}
	}
	for key, col := range e.prefCols { // nodeset.go:69-82: exact values
		if v, ok := label(key); ok {
			attr(col, e.prefID(v))
		}
	}
	if _, in := e.volIn[n.ID]; e.volCol != C.PE_NONE && in {
		attr(e.volCol, volMark)
	}
	r.attr_cnt = C.uint32_t(len(b.attrs)) - r.attr_off
	r.gen_off = C.uint32_t(len(b.gens))
	if ni.AvailableResources != nil {
		seen := map[string]bool{}
		for _, g := range ni.AvailableResources.Generic {
			kind := genericresource.Kind(g)
			if seen[kind] {
				continue
			}
			seen[kind] = true
			var cell int64
			if dr := g.GetDiscreteResourceSpec(); dr != nil { // validate.go:36-48
				cell = int64(C.PE_GEN_DISCRETE) | dr.Value<<2
			} else {
				cell = int64(C.PE_GEN_NAMED) | int64(len(genericresource.GetResource(kind, ni.AvailableResources.Generic)))<<2
			}
			b.gens = append(b.gens, C.pe_kv64{key: C.uint32_t(intern(e.kinds, kind)), value: C.int64_t(cell)})
		}
	}
	r.gen_cnt = C.uint32_t(len(b.gens)) - r.gen_off
	r.svc_off = C.uint32_t(len(b.svcs))
	for svc, c := range ni.ActiveTasksCountByService {
		if c != 0 {
			b.svcs = append(b.svcs, C.pe_kv32{key: C.uint32_t(e.serviceID(svc)), value: C.uint32_t(c)})
		}
	}
	r.svc_cnt = C.uint32_t(len(b.svcs)) - r.svc_off
	r.port_off = C.uint32_t(len(b.ports))
	for p := range ni.usedHostPorts {
		b.ports = append(b.ports, C.uint32_t(e.portSlot(p)))
	}
	r.port_cnt = C.uint32_t(len(b.ports)) - r.port_off
	r.plug_off = C.uint32_t(len(b.plugs))
	if d != nil && d.Engine != nil {
		for k, slot := range e.plugins { // filter.go:186-205
			for _, p := range d.Engine.Plugins {
				if p.Type == k[0] && (p.Name == k[1] || p.Name == k[1]+":latest") {
					b.plugs = append(b.plugs, C.uint32_t(slot))
					break
				}
			}
		}
		for _, p := range d.Engine.Plugins {
			if p.Type == "Log" {
				r.flags |= C.PE_NODE_HAS_LOGPLUGIN
				break
			}
		}
	}
	r.plug_cnt = C.uint32_t(len(b.plugs)) - r.plug_off
	b.rows = append(b.rows, r)
}

func ptr[T any](s []T) *T {
	if len(s) == 0 {
		return nil
	}
	return &s[0]
}

// flushRows uploads what store events changed since the last engine call (createOrUpdateNode scheduler.go:368-396,
// nodeSet.remove nodeset.go:46-48): all rows when membership or a dictionary changed, the touched rows otherwise.
func (e *gpuEngine) flushRows(ns *nodeSet) error {
	if !e.layout && len(e.dirty) == 0 {
		return nil
	}
	var b rowBatch
	if e.layout {
		e.order = e.order[:0]
		for id := range ns.nodes {
			e.order = append(e.order, id)
		}
		sort.Strings(e.order)
		e.index = make(map[string]uint32, len(e.order))
		for i, id := range e.order {
			e.index[id] = uint32(i)
			ni := ns.nodes[id]
			e.encodeRow(uint32(i), &ni, &b)
		}
		if rc := C.pe_set_node_count(e.h, C.uint32_t(len(e.order))); rc != C.PE_OK {
			return e.err("pe_set_node_count")
		}
	} else {
		for id := range e.dirty {
			if ni, ok := ns.nodes[id]; ok {
				e.encodeRow(e.index[id], &ni, &b)
			}
		}
	}
	e.layout, e.dirty = false, map[string]struct{}{}
	if len(b.rows) == 0 {
		return nil
	}
	if rc := C.pe_node_upsert(e.h, ptr(b.rows), C.uint32_t(len(b.rows)), ptr(b.attrs), ptr(b.gens), ptr(b.svcs), ptr(b.ports), ptr(b.plugs)); rc != C.PE_OK {
		return e.err("pe_node_upsert")
	}
	return nil
}

// nodeChanged / nodeSetChanged are called from createOrUpdateNode, nodeSet.remove, and wherever addTask/removeTask/
// taskFailed touch a NodeInfo outside a tick (createTask, updateTask, deleteTask, commit rollback).
func (e *gpuEngine) nodeChanged(id string) { e.dirty[id] = struct{}{} }
func (e *gpuEngine) nodeSetChanged()      { e.layout = true }

// ---- group descriptors: the SetTask methods of filter.go ----------------------------------------------------------

type tickBuf struct {
	groups []C.pe_group
	flags  []C.uint8_t
	gens   []C.pe_generic_want
	cons   []C.pe_constraint
	ips    []C.pe_ip_constraint
	plats  []C.pe_platform
	ports  []C.uint32_t
	plugs  []C.uint32_t
	fails  []C.pe_node_fail
	failAt map[versionedService][2]uint32
}

func (b *tickBuf) view() C.pe_tick {
	return C.pe_tick{groups: ptr(b.groups), n_groups: C.uint32_t(len(b.groups)), task_flags: ptr(b.flags), n_tasks: C.uint32_t(len(b.flags)),
		gens: ptr(b.gens), n_gens: C.uint32_t(len(b.gens)), cons: ptr(b.cons), n_cons: C.uint32_t(len(b.cons)),
		ips: ptr(b.ips), n_ips: C.uint32_t(len(b.ips)), plats: ptr(b.plats), n_plats: C.uint32_t(len(b.plats)),
		ports: ptr(b.ports), n_ports: C.uint32_t(len(b.ports)), plugs: ptr(b.plugs), n_plugs: C.uint32_t(len(b.plugs)),
		fails: ptr(b.fails), n_fails: C.uint32_t(len(b.fails))}
}

// buildFailLists: countRecentFailures(now) for the nodes that have entries, one sorted list per (service, version),
// built once per tick (scheduler.go:711-712 evaluates it per comparison).
func (s *Scheduler) buildFailLists(b *tickBuf, now time.Time) {
	per := map[versionedService][]C.pe_node_fail{}
	for i, id := range s.gpu.order {
		ni := s.nodeSet.nodes[id]
		for vs := range ni.recentFailures {
			t := &api.Task{ServiceID: vs.serviceID, SpecVersion: &vs.specVersion}
			if c := ni.countRecentFailures(now, t); c > 0 {
				per[vs] = append(per[vs], C.pe_node_fail{node_idx: C.uint32_t(i), count: C.uint32_t(c)})
			}
		}
	}
	b.failAt = map[versionedService][2]uint32{}
	for vs, l := range per {
		b.failAt[vs] = [2]uint32{uint32(len(b.fails)), uint32(len(l))}
		b.fails = append(b.fails, l...)
	}
}

// compileConstraints: constraint.Parse (constraint.go:40-81) + the key dispatch of NodeMatches (:107-207) as integer
// programs.  false = Parse failed: the filter stays disabled (filter.go:229-236).
func (e *gpuEngine) compileConstraints(exprs []string, g *C.pe_group, b *tickBuf) bool {
	cs, perr := constraint.Parse(exprs)
	if perr != nil {
		return false
	}
	var cons []C.pe_constraint
	var ips []C.pe_ip_constraint
	never := false
	for _, c := range cs {
		neq := C.uint32_t(0)
		if c.Operator() != constraint.EqualOperator() {
			neq = 1
		}
		key := c.Key()
		col := uint32(C.PE_NONE)
		switch {
		case strings.EqualFold(key, constraint.NodeIDKey):
			col = C.PE_ATTR_NODE_ID
		case strings.EqualFold(key, constraint.NodeHostnameKey):
			col = C.PE_ATTR_HOSTNAME
		case strings.EqualFold(key, constraint.NodeRoleKey):
			col = C.PE_ATTR_ROLE
		case strings.EqualFold(key, constraint.NodePlatformOSKey):
			col = C.PE_ATTR_OS
		case strings.EqualFold(key, constraint.NodePlatformArchKey):
			col = C.PE_ATTR_ARCH
		case strings.EqualFold(key, constraint.NodeIPKey):
			ic := C.pe_ip_constraint{neq: neq}
			if ip := net.ParseIP(c.Exp()); ip != nil {
				fillIP(&ic, ip.To16(), net.CIDRMask(128, 128), false, false)
				ips = append(ips, ic)
			} else if _, subnet, cerr := net.ParseCIDR(c.Exp()); cerr == nil {
				v4 := subnet.IP.To4() != nil
				ones, _ := subnet.Mask.Size()
				if v4 {
					ones += 96
				}
				fillIP(&ic, subnet.IP.To16(), net.CIDRMask(ones, 128), true, v4)
				ips = append(ips, ic)
			} else {
				never = true // constraint.go:144-146
			}
			continue
		case len(key) > len(constraint.NodeLabelPrefix) && strings.EqualFold(key[:len(constraint.NodeLabelPrefix)], constraint.NodeLabelPrefix):
			col = e.column(e.labelCols, "n:"+key[len(constraint.NodeLabelPrefix):])
		case len(key) > len(constraint.EngineLabelPrefix) && strings.EqualFold(key[:len(constraint.EngineLabelPrefix)], constraint.EngineLabelPrefix):
			col = e.column(e.labelCols, "e:"+key[len(constraint.EngineLabelPrefix):])
		default:
			never = true // unknown key rejects every node, constraint.go:200-203
			continue
		}
		cons = append(cons, C.pe_constraint{col: C.uint32_t(col), value: C.uint32_t(e.valueID(c.Exp())), neq: neq})
	}
	g.con_off, g.con_cnt = C.uint32_t(len(b.cons)), C.uint32_t(len(cons))
	b.cons = append(b.cons, cons...)
	g.ip_off, g.ip_cnt = C.uint32_t(len(b.ips)), C.uint32_t(len(ips))
	b.ips = append(b.ips, ips...)
	if never {
		g.flags |= C.PE_G_CONSTRAINT_NEVER
	}
	return true
}

func fillIP(ic *C.pe_ip_constraint, ip16 net.IP, mask net.IPMask, cidr, v4 bool) {
	for w := 0; w < 4; w++ {
		m := uint32(mask[4*w])<<24 | uint32(mask[4*w+1])<<16 | uint32(mask[4*w+2])<<8 | uint32(mask[4*w+3])
		a := uint32(ip16[4*w])<<24 | uint32(ip16[4*w+1])<<16 | uint32(ip16[4*w+2])<<8 | uint32(ip16[4*w+3])
		ic.mask[w], ic.net[w] = C.uint32_t(m), C.uint32_t(a&m)
	}
	if cidr {
		ic.is_cidr = 1
	}
	if v4 {
		ic.is_v4 = 1
	}
}

// encodeGroup: one pe_group for tasks that share a spec (the SetTask methods: filter.go:35,60,118,224,259,328,369).
func (s *Scheduler) encodeGroup(tasks []*api.Task, b *tickBuf, now time.Time) error {
	e, t := s.gpu, tasks[0]
	g := C.pe_group{log_plugin: C.PE_NONE, svc_id: C.uint32_t(e.serviceID(t.ServiceID)), n_tasks: C.uint32_t(len(tasks)),
		task_off: C.uint32_t(len(b.flags)), filter_mask: 1 << C.PE_F_READY}
	for _, x := range tasks {
		f := C.uint8_t(0)
		if x.DesiredState <= api.TaskStateCompleted { // nodeinfo.go:148
			f = C.PE_T_COUNTS
		}
		b.flags = append(b.flags, f)
	}
	// (cluster mounts: VolumesFilter is evaluated by the shim, which hands the engine the node set: scheduleVolumeGroup)
	res := taskReservations(t.Spec) // nodeinfo.go:156-161
	g.cpu_res, g.mem_res = C.int64_t(res.NanoCPUs), C.int64_t(res.MemoryBytes)
	g.gen_off = C.uint32_t(len(b.gens))
	for _, w := range res.Generic {
		dr := w.GetDiscreteResourceSpec()
		if dr == nil {
			return errors.New("task reserves a named generic resource (validate.go:26-29 rejects it too)")
		}
		b.gens = append(b.gens, C.pe_generic_want{kind: C.uint32_t(intern(e.kinds, dr.Kind)), value: C.int64_t(dr.Value)})
	}
	g.gen_cnt = C.uint32_t(len(b.gens)) - g.gen_off
	if r := t.Spec.Resources; r != nil && r.Reservations != nil && (res.NanoCPUs != 0 || res.MemoryBytes != 0 || len(res.Generic) != 0) {
		g.filter_mask |= 1 << C.PE_F_RESOURCE // filter.go:60-73
	}
	// PluginFilter.SetTask, filter.go:118-139
	g.plug_off = C.uint32_t(len(b.plugs))
	plugin := false
	if c := t.Spec.GetContainer(); c != nil {
		for _, m := range c.Mounts {
			if m.Type == api.MountTypeVolume && m.VolumeOptions != nil && m.VolumeOptions.DriverConfig != nil &&
				m.VolumeOptions.DriverConfig.Name != "" && m.VolumeOptions.DriverConfig.Name != "local" {
				b.plugs = append(b.plugs, C.uint32_t(e.pluginSlot("Volume", m.VolumeOptions.DriverConfig.Name)))
				plugin = true
			}
		}
	}
	for _, na := range t.Networks {
		plugin = true
		if na.Network != nil && na.Network.DriverState != nil && na.Network.DriverState.Name != "" {
			b.plugs = append(b.plugs, C.uint32_t(e.pluginSlot("Network", na.Network.DriverState.Name)))
		}
	}
	if t.Spec.LogDriver != nil {
		plugin = true
		if n := t.Spec.LogDriver.Name; n != "none" && n != "" {
			g.flags |= C.PE_G_LOG_DRIVER
			g.log_plugin = C.uint32_t(e.pluginSlot("Log", n))
		}
	}
	g.plug_cnt = C.uint32_t(len(b.plugs)) - g.plug_off
	if plugin {
		g.filter_mask |= 1 << C.PE_F_PLUGIN
	}
	if p := t.Spec.Placement; p != nil {
		if len(p.Constraints) != 0 && e.compileConstraints(p.Constraints, &g, b) { // filter.go:224-238
			g.filter_mask |= 1 << C.PE_F_CONSTRAINT
		}
		if len(p.Platforms) != 0 { // filter.go:259-269
			g.filter_mask |= 1 << C.PE_F_PLATFORM
			g.plat_off = C.uint32_t(len(b.plats))
			for _, pf := range p.Platforms {
				var pl C.pe_platform // 0 = wildcard
				if pf.OS != "" {
					pl.os_id = C.uint32_t(intern(e.exact, pf.OS))
				}
				if pf.Architecture != "" {
					pl.arch_id = C.uint32_t(intern(e.exact, normArch(pf.Architecture)))
				}
				b.plats = append(b.plats, pl)
			}
			g.plat_cnt = C.uint32_t(len(b.plats)) - g.plat_off
		}
		if p.MaxReplicas > 0 { // filter.go:369-376
			g.filter_mask |= 1 << C.PE_F_MAXREPLICAS
			g.max_replicas = C.uint64_t(p.MaxReplicas)
		}
	}
	g.port_off = C.uint32_t(len(b.ports))
	if t.Endpoint != nil { // filter.go:328-339 reads the allocator-filled Endpoint.Ports
		for _, p := range t.Endpoint.Ports {
			if p.PublishMode == api.PublishModeHost && p.PublishedPort != 0 {
				b.ports = append(b.ports, C.uint32_t(e.portSlot(hostPortSpec{protocol: p.Protocol, publishedPort: p.PublishedPort})))
			}
		}
	}
	g.port_cnt = C.uint32_t(len(b.ports)) - g.port_off
	if g.port_cnt != 0 {
		g.filter_mask |= 1 << C.PE_F_HOSTPORT
	}
	if b.failAt == nil {
		s.buildFailLists(b, now)
	}
	vs := versionedService{serviceID: t.ServiceID}
	if t.SpecVersion != nil {
		vs.specVersion = *t.SpecVersion
	}
	if r, ok := b.failAt[vs]; ok {
		g.fail_off, g.fail_cnt = C.uint32_t(r[0]), C.uint32_t(r[1])
	}
	b.groups = append(b.groups, g)
	return nil
}

// ---- the group loop of tick (scheduler.go:464-469) -----------------------------------------------------------------

func sortedTasks(group map[string]*api.Task) []*api.Task {
	ts := make([]*api.Task, 0, len(group))
	for _, t := range group {
		ts = append(ts, t)
	}
	sort.Slice(ts, func(i, j int) bool { return ts[i].ID < ts[j].ID }) // canonical order inside a group
	return ts
}

// assign is the body of scheduleNTasksOnNodes that stays on the host (scheduler.go:871-893).
func (s *Scheduler) assign(t *api.Task, nodeID string, group map[string]*api.Task, decisions map[string]schedulingDecision) {
	newT := *t
	if hasClusterMounts(t) { // scheduler.go:857-874 (an error is only logged there)
		if ni, err := s.nodeSet.nodeInfo(nodeID); err == nil {
			newT.Volumes, _ = s.volumes.chooseTaskVolumes(t, &ni)
		}
	}
	newT.NodeID = nodeID
	s.volumes.reserveTaskVolumes(&newT)
	newT.Status = api.TaskStatus{State: api.TaskStateAssigned, Timestamp: ptypes.MustTimestampProto(time.Now()),
		Message: "scheduler assigned task to node"}
	s.allTasks[t.ID] = &newT
	if ni, err := s.nodeSet.nodeInfo(nodeID); err == nil && ni.addTask(&newT) { // host mirror incl. named generic members
		s.nodeSet.updateNode(ni) // (the device row already moved: not marked dirty)
	}
	decisions[t.ID] = schedulingDecision{old: t, new: &newT}
	delete(group, t.ID)
}

// scheduleTickGPU replaces the two loops over scheduleTaskGroup: groups in canonical order (ascending (ServiceID,
// SpecVersion.Index), then one-offs by task ID); maximal runs of groups without placement preferences go to the engine
// in ONE pe_schedule call, a group with preferences is walked leaf by leaf in between, a group with cluster mounts is
// restricted to the nodes VolumesFilter allows.
func (s *Scheduler) scheduleTickGPU(ctx context.Context, groups []map[string]*api.Task, decisions map[string]schedulingDecision) {
	now := time.Now() // sampled once per tick (the reference samples it once per group, scheduler.go:706)
	s.recycleServiceIDs()
	var run []map[string]*api.Task
	flush := func() {
		if len(run) != 0 {
			s.scheduleRunGPU(ctx, run, decisions, now)
			run = run[:0]
		}
	}
	for _, g := range groups {
		var any *api.Task
		for _, t := range g {
			any = t
			break
		}
		if any != nil && hasClusterMounts(any) {
			flush()
			s.scheduleVolumeGroup(ctx, g, decisions, now)
			continue
		}
		if any != nil && len(s.preferenceLevels(any)) != 0 {
			flush()
			s.schedulePreferenceGroup(ctx, g, decisions, now)
			continue
		}
		run = append(run, g)
	}
	flush()
}

func (s *Scheduler) scheduleRunGPU(ctx context.Context, groups []map[string]*api.Task, decisions map[string]schedulingDecision, now time.Time) {
	var b tickBuf
	var taken [][]*api.Task
	var of []map[string]*api.Task
	for _, g := range groups {
		ts := sortedTasks(g)
		probe := b
		if err := s.encodeGroup(ts, &probe, now); err != nil { // its tasks stay pending with the reason, like noSuitableNode
			s.noSuitableNodeWith(ctx, g, decisions, "unsupported by the placement engine: "+err.Error())
			continue
		}
		b = probe
		taken, of = append(taken, ts), append(of, g)
	}
	if len(taken) == 0 {
		return
	}
	requeue := func(err error) {
		log.G(ctx).WithError(err).Error("placement engine")
		for _, g := range of {
			for _, t := range g {
				s.enqueue(t)
			}
		}
		s.gpu.nodeSetChanged() // the device rows may have moved part-way: rebuild them from NodeInfo
	}
	if err := s.gpu.flushRows(&s.nodeSet); err != nil {
		requeue(err)
		return
	}
	outNode := make([]C.uint32_t, len(b.flags))
	outFail := make([]C.uint32_t, len(b.groups)*C.PE_NUM_FILTERS)
	tick := b.view() // flat arrays; the engine keeps no pointer after the call returns
	if rc := C.pe_schedule(s.gpu.h, &tick, ptr(outNode), ptr(outFail)); rc != C.PE_OK {
		requeue(s.gpu.err("pe_schedule"))
		return
	}
	for gi, ts := range taken {
		off := int(b.groups[gi].task_off)
		for i, t := range ts {
			if idx := uint32(outNode[off+i]); idx != C.PE_NONE {
				s.assign(t, s.gpu.order[idx], of[gi], decisions)
			}
		}
		if len(of[gi]) != 0 {
			s.noSuitableNodeWith(ctx, of[gi], decisions, explainCounters(outFail[gi*C.PE_NUM_FILTERS:]))
		}
	}
}

// ---- cluster (CSI) volumes: VolumesFilter (filter.go:388-447) is evaluated HERE against s.volumes; the engine gets the
// answer as the group's node set -- the rows marked in one attribute column, named by a leaf term -- and does the rest
// (DESIGN.md 4.7): ONE engine group when no placement can change volume availability, the reference's fill loop with one
// engine question per step otherwise (scheduleVolumeGroupStepwise).

func hasClusterMounts(t *api.Task) bool {
	if c := t.Spec.GetContainer(); c != nil {
		for _, m := range c.Mounts {
			if m.Type == api.MountTypeCluster {
				return true
			}
		}
	}
	return false
}

// volumesStaticFor: can a placement of one task of this spec change what VolumesFilter answers for the next one?
// (scope SINGLE_NODE pins the volume to its first node; sharing NONE and ONE_WRITER writers count users: volumes.go:257-318)
func (s *Scheduler) volumesStaticFor(t *api.Task) bool {
	for _, m := range t.Spec.GetContainer().Mounts {
		if m.Type != api.MountTypeCluster {
			continue
		}
		var ids []string
		if group := strings.TrimPrefix(m.Source, "group:"); group != m.Source {
			for id := range s.volumes.byGroup[group] {
				ids = append(ids, id)
			}
		} else if id, ok := s.volumes.byName[m.Source]; ok {
			ids = append(ids, id)
		}
		for _, id := range ids {
			am := s.volumes.volumes[id].volume.Spec.AccessMode
			if am.Scope != api.VolumeScopeMultiNode || am.Sharing == api.VolumeSharingNone ||
				(am.Sharing == api.VolumeSharingOneWriter && !m.ReadOnly) {
				return false
			}
		}
	}
	return true
}

func (e *gpuEngine) markVolumeNodes(ids []string) {
	if e.volCol == C.PE_NONE {
		e.volCol = e.nextCol // a column of its own, filled for every node on the next upload
		e.nextCol++
		e.nodeSetChanged()
	}
	// only the rows whose membership changes are uploaded: consecutive groups with the same mounts mostly ask for the
	// set that is already marked
	next := make(map[string]struct{}, len(ids))
	for _, id := range ids {
		next[id] = struct{}{}
	}
	for id := range e.volIn {
		if _, ok := next[id]; !ok {
			delete(e.volIn, id)
			e.nodeChanged(id)
		}
	}
	for _, id := range ids {
		if _, ok := e.volIn[id]; !ok {
			e.volIn[id] = struct{}{}
			e.nodeChanged(id)
		}
	}
}

// countExcluded: the first failing filter of every node of ids, none of which passes VolumesFilter: a device filter if one
// fails, VolumesFilter (the last of the pipeline) otherwise.
func (s *Scheduler) countExcluded(t *api.Task, ids []string, cnt []C.uint32_t, now time.Time) error {
	ok, fail, err := s.fitMany(t, ids, true, now)
	if err != nil {
		return err
	}
	for i := range ids {
		switch ok[i] {
		case 1:
			cnt[C.PE_F_VOLUMES]++
		case 0:
			for f := 0; f < C.PE_NUM_FILTERS; f++ {
				cnt[f] += fail[i*C.PE_NUM_FILTERS+f]
			}
		}
	}
	return nil
}

// fitMany asks taskFitNode's question for one task of this spec on each of ids (batches of pe_fit requests).  A success
// RESERVES on the device; with undo the reservations are taken back (the rows are uploaded again before the next call).
func (s *Scheduler) fitMany(t *api.Task, ids []string, undo bool, now time.Time) (ok []C.uint8_t, fail []C.uint32_t, err error) {
	const batch = 4096
	ok, fail = make([]C.uint8_t, len(ids)), make([]C.uint32_t, len(ids)*C.PE_NUM_FILTERS)
	for lo := 0; lo < len(ids); lo += batch {
		part := ids[lo:min(lo+batch, len(ids))]
		var b tickBuf
		idx := make([]C.uint32_t, len(part))
		for i, id := range part {
			if err = s.encodeGroup([]*api.Task{t}, &b, now); err != nil {
				return
			}
			idx[i] = C.uint32_t(s.gpu.index[id])
		}
		if err = s.gpu.flushRows(&s.nodeSet); err != nil {
			return
		}
		tick := b.view()
		if rc := C.pe_fit(s.gpu.h, &tick, ptr(idx), &ok[lo], &fail[lo*C.PE_NUM_FILTERS]); rc != C.PE_OK {
			err = s.gpu.err("pe_fit")
			return
		}
		if undo {
			for i, id := range part {
				if ok[lo+i] == 1 {
					s.gpu.nodeChanged(id)
				}
			}
		}
	}
	return
}

// rankKey is nodeLess (scheduler.go:708-734) plus the canonical tie-break, as a key.
type rankKey struct {
	f, s, a int
	id      string
}

func (a rankKey) lessNoTie(b rankKey) bool {
	if a.f != b.f {
		return a.f < b.f
	}
	if a.s != b.s {
		return a.s < b.s
	}
	return a.a < b.a
}
func (a rankKey) less(b rankKey) bool {
	if a.lessNoTie(b) {
		return true
	}
	if b.lessNoTie(a) {
		return false
	}
	return a.id < b.id
}

// rankOf is the rank key of a node for tasks of t's spec.
func (s *Scheduler) rankOf(id string, t *api.Task, now time.Time) rankKey {
	ni, _ := s.nodeSet.nodeInfo(id)
	fl := ni.countRecentFailures(now, t)
	if fl < maxFailures {
		fl = 0
	}
	return rankKey{fl, ni.ActiveTasksCountByService[t.ServiceID], ni.ActiveTasksCount, id}
}

// probeBest: the `want` best feasible nodes of (leaf terms AND the volume set), in rank order: successive one-task groups,
// each answer taken out of the volume set and its reservation taken back (the row goes up again before the next call).
// firstFail: the counters of the first question when it found no node.  errUnsupported wraps an encoder refusal.
func (s *Scheduler) probeBest(t *api.Task, leafTerms []C.pe_constraint, want int, now time.Time) (cand []string, firstFail []C.uint32_t, err error) {
	e := s.gpu
	var b tickBuf
	if err = s.encodeGroup([]*api.Task{t}, &b, now); err != nil {
		return nil, nil, fmt.Errorf("%w: %v", errUnsupported, err)
	}
	g := &b.groups[len(b.groups)-1]
	if g.con_cnt == 0 {
		g.con_off = C.uint32_t(len(b.cons))
	}
	b.cons = append(b.cons, leafTerms...)
	b.cons = append(b.cons, C.pe_constraint{col: C.uint32_t(e.volCol), value: C.uint32_t(volMark)})
	g.leaf_cnt = C.uint32_t(len(leafTerms) + 1)
	for i := 0; i < want; i++ {
		if err = e.flushRows(&s.nodeSet); err != nil {
			return
		}
		var pn C.uint32_t = C.PE_NONE
		pf := make([]C.uint32_t, C.PE_NUM_FILTERS)
		tick := b.view()
		if rc := C.pe_schedule(e.h, &tick, &pn, ptr(pf)); rc != C.PE_OK {
			return nil, nil, e.err("pe_schedule")
		}
		if pn == C.PE_NONE {
			if i == 0 {
				firstFail = pf
			}
			break
		}
		id := e.order[pn]
		cand = append(cand, id)
		delete(e.volIn, id)
		e.nodeChanged(id)
	}
	return
}

var errUnsupported = errors.New("unsupported by the placement engine")

// fillStepwise is scheduleNTasksOnNodes (scheduler.go:844-924) over an ordered node list with one engine question per step:
// "does this node still pass, and if so the next task goes there" = pe_fit (device filters + reservation), then
// VolumesFilter here -- the pipeline's order, so cnt moves as the reference's failure counters do (cleared by every pass).
// Returns the tasks placed (a prefix of ts) and whether a re-check passed.
func (s *Scheduler) fillStepwise(ts []*api.Task, cand []string, cnt []C.uint32_t, group map[string]*api.Task,
	decisions map[string]schedulingDecision, now time.Time) (placed int, passed bool, err error) {
	k, m := len(ts), len(cand)
	if k == 0 || m == 0 {
		return
	}
	f := &VolumesFilter{vs: s.volumes}
	failed := make([]bool, m)
	it := 0
	if ok, _, ferr := s.fitMany(ts[0], cand[:1], false, now); ferr != nil {
		return 0, false, ferr
	} else if ok[0] != 1 {
		return 0, false, errors.New("placement engine: the best node of a group does not fit its first task")
	}
fill:
	for {
		nid := cand[it%m]
		s.assign(ts[placed], nid, group, decisions)
		placed++
		if placed == k {
			return
		}
		if it+1 < m {
			if s.rankOf(cand[(it+1)%m], ts[0], now).lessNoTie(s.rankOf(nid, ts[0], now)) { // first pass: level the nodes
				it++
			}
		} else {
			it++ // later passes: one task per node
		}
		for orig := it; ; {
			if i := it % m; !failed[i] {
				ok, fail, ferr := s.fitMany(ts[placed], cand[i:i+1], false, now)
				if ferr != nil {
					return placed, passed, ferr
				}
				ni, _ := s.nodeSet.nodeInfo(cand[i])
				f.SetTask(ts[placed])
				if ok[0] == 1 && f.Check(&ni) {
					for j := range cnt {
						cnt[j] = 0 // pipeline.go:64-66
					}
					passed = true
					continue fill
				}
				if ok[0] == 1 {
					cnt[C.PE_F_VOLUMES]++
					s.gpu.nodeChanged(cand[i]) // the reservation is taken back
				} else {
					for j := range cnt {
						cnt[j] += fail[j]
					}
				}
				failed[i] = true
			}
			it++
			if it-orig == m {
				return // none of the nodes meets the constraints any more
			}
		}
	}
}

// scheduleVolumeGroupStepwise: a group whose volume availability moves with every placement.  The reference re-runs the
// whole pipeline, VolumesFilter included, on the next node before every further placement (scheduler.go:912-920); this
// walks that loop and asks the engine one question per step (DESIGN.md 4.7).  s.gpu.markVolumeNodes(allowed) was called.
func (s *Scheduler) scheduleVolumeGroupStepwise(ctx context.Context, ts []*api.Task, group map[string]*api.Task, allowedIDs, excluded []string,
	decisions map[string]schedulingDecision, now time.Time) error {
	e, t, k := s.gpu, ts[0], len(ts)
	allowed := make(map[string]struct{}, len(allowedIDs))
	for _, id := range allowedIDs {
		allowed[id] = struct{}{}
	}
	cnt := make([]C.uint32_t, C.PE_NUM_FILTERS)
	cand, firstFail, err := s.probeBest(t, nil, k, now)
	if errors.Is(err, errUnsupported) {
		s.noSuitableNodeWith(ctx, group, decisions, err.Error())
		return nil
	} else if err != nil {
		return err
	}
	if len(cand) == 0 { // nothing passed: every node is counted by its first failing filter
		if firstFail != nil {
			cnt = firstFail
		}
		if err := s.countExcluded(t, excluded, cnt, now); err != nil {
			return err
		}
		s.noSuitableNodeWith(ctx, group, decisions, explainCounters(cnt))
		return nil
	}
	firstKey := s.rankOf(cand[0], t, now)
	placed, passed, err := s.fillStepwise(ts, cand, cnt, group, decisions, now)
	if err != nil {
		return err
	}
	if placed == k {
		return nil
	}
	if !passed {
		// no re-check passed: the tree building's counters still stand.  Replay its walk (ascending node ID): the pipeline
		// ran on every node while the heap had room, afterwards only on nodes ranking ahead of the heap's worst
		// (nodeset.go:103-121); a pass clears the counters.
		var ids []string
		for _, id := range e.order {
			if id != cand[0] {
				ids = append(ids, id)
			}
		}
		ok, fail, err := s.fitMany(ts[placed], ids, true, now)
		if err != nil {
			return err
		}
		tree := make([]C.uint32_t, C.PE_NUM_FILTERS)
		var heap []rankKey
		worst := func() int {
			w := 0
			for i := range heap {
				if heap[w].less(heap[i]) {
					w = i
				}
			}
			return w
		}
		j := 0
		for _, id := range e.order {
			first := id == cand[0]
			rk, q := firstKey, 0
			if !first {
				rk, q = s.rankOf(id, t, now), j
				j++
			}
			w := -1
			if len(heap) >= k {
				if w = worst(); !rk.less(heap[w]) {
					continue // the pipeline was not run on it
				}
			}
			_, in := allowed[id]
			if !(first || (ok[q] == 1 && in)) {
				if ok[q] == 1 {
					tree[C.PE_F_VOLUMES]++
				} else if ok[q] == 0 {
					for x := range tree {
						tree[x] += fail[q*C.PE_NUM_FILTERS+x]
					}
				}
				continue
			}
			for x := range tree {
				tree[x] = 0
			}
			if w >= 0 {
				heap[w] = rk
			} else {
				heap = append(heap, rk)
			}
		}
		for x := range cnt {
			cnt[x] += tree[x]
		}
	}
	s.noSuitableNodeWith(ctx, group, decisions, explainCounters(cnt))
	return nil
}

func (s *Scheduler) scheduleVolumeGroup(ctx context.Context, group map[string]*api.Task, decisions map[string]schedulingDecision, now time.Time) {
	ts := sortedTasks(group)
	t := ts[0]
	refuse := func(why string) {
		s.noSuitableNodeWith(ctx, group, decisions, "unsupported by the placement engine: "+why)
	}
	prefs := len(s.preferenceLevels(t)) != 0
	requeue := func(err error) {
		log.G(ctx).WithError(err).Error("placement engine")
		for _, x := range group {
			s.enqueue(x)
		}
		s.gpu.nodeSetChanged()
	}
	// (the C++ shim also keeps this answer between groups that ask the same question -- same mounts, no volume, counted use
	// or node changed since: VolumeBook::epoch / node_epoch in scheduler_host.cpp; here that needs a change counter in
	// volumeSet.reserveVolume / releaseVolume / addOrUpdateVolume, three one-line additions to volumes.go)
	f := &VolumesFilter{vs: s.volumes}
	f.SetTask(t)
	var allowed, excluded []string
	for _, id := range s.gpu.order { // canonical order: ascending node ID
		ni, err := s.nodeSet.nodeInfo(id)
		if err != nil {
			continue
		}
		if f.Check(&ni) {
			allowed = append(allowed, id)
		} else {
			excluded = append(excluded, id)
		}
	}
	s.gpu.markVolumeNodes(allowed)
	if prefs {
		// the tree's branches and task sums do not depend on the pipeline (nodeset.go:59-101): the preference walk runs
		// unchanged and every leaf visit carries the volume term next to the leaf's own terms (fillLeaf)
		s.gpu.leafAlsoInVolumeSet = true
		s.gpu.leafStepwise = len(ts) > 1 && !s.volumesStaticFor(t) // the walk of the leaf's heap, step by step (fillLeafStepwise)
		s.schedulePreferenceGroup(ctx, group, decisions, now)
		s.gpu.leafAlsoInVolumeSet, s.gpu.leafStepwise = false, false
		return
	}
	if len(ts) > 1 && !s.volumesStaticFor(t) {
		if err := s.scheduleVolumeGroupStepwise(ctx, ts, group, allowed, excluded, decisions, now); err != nil {
			requeue(err)
		}
		return
	}
	var b tickBuf
	if err := s.encodeGroup(ts, &b, now); err != nil {
		refuse(err.Error())
		return
	}
	g := &b.groups[len(b.groups)-1]
	if g.con_cnt == 0 {
		g.con_off = C.uint32_t(len(b.cons)) // the leaf term follows the group's constraints
	}
	b.cons = append(b.cons, C.pe_constraint{col: C.uint32_t(s.gpu.volCol), value: C.uint32_t(volMark)})
	g.leaf_cnt = 1
	if err := s.gpu.flushRows(&s.nodeSet); err != nil {
		requeue(err)
		return
	}
	outNode := make([]C.uint32_t, len(ts))
	cnt := make([]C.uint32_t, C.PE_NUM_FILTERS)
	tick := b.view()
	if rc := C.pe_schedule(s.gpu.h, &tick, ptr(outNode), ptr(cnt)); rc != C.PE_OK {
		requeue(s.gpu.err("pe_schedule"))
		return
	}
	placed, firstPlaced := 0, ""
	for i, x := range ts {
		if idx := uint32(outNode[i]); idx != C.PE_NONE {
			if placed == 0 {
				firstPlaced = s.gpu.order[idx]
			}
			s.assign(x, s.gpu.order[idx], group, decisions) // chooses and reserves the attachments
			placed++
		}
	}
	if len(group) == 0 {
		return
	}
	// Explain() reports what failed after the last node that PASSED (pipeline.go:55-68): every excluded node when nothing was
	// placed; the excluded nodes after the one node that passed when exactly one task was; none when a re-check inside
	// the fill loop passed (two or more placed).  DESIGN.md 4.7.
	var counted []string
	switch placed {
	case 0:
		counted = excluded
	case 1:
		for _, id := range excluded {
			if id > firstPlaced {
				counted = append(counted, id)
			}
		}
	}
	if len(counted) != 0 {
		if err := s.countExcluded(t, counted, cnt, now); err != nil {
			requeue(err)
			return
		}
	}
	s.noSuitableNodeWith(ctx, group, decisions, explainCounters(cnt))
}

// ---- placement preferences: nodeSet.tree's branches (nodeset.go:59-101) + scheduleNTasksOnSubtree (scheduler.go:772-825)

func (s *Scheduler) preferenceLevels(t *api.Task) []uint32 {
	var cols []uint32
	if t.Spec.Placement == nil {
		return nil
	}
	for _, pref := range t.Spec.Placement.Preferences {
		spread := pref.GetSpread()
		if spread == nil {
			continue
		}
		d := spread.SpreadDescriptor
		switch {
		case len(d) > len(constraint.NodeLabelPrefix) && strings.EqualFold(d[:len(constraint.NodeLabelPrefix)], constraint.NodeLabelPrefix):
			cols = append(cols, s.gpu.column(s.gpu.prefCols, "n:"+d[len(constraint.NodeLabelPrefix):]))
		case len(d) > len(constraint.EngineLabelPrefix) && strings.EqualFold(d[:len(constraint.EngineLabelPrefix)], constraint.EngineLabelPrefix):
			cols = append(cols, s.gpu.column(s.gpu.prefCols, "e:"+d[len(constraint.EngineLabelPrefix):]))
		}
	}
	return cols
}

type prefTree struct {
	tasks int
	next  map[string]*prefTree
	leaf  []C.pe_constraint // the (column == value) path that names a leaf
	cand  []string          // fillLeafStepwise: the leaf's heap (decision_tree.go), best first
	built bool
}

type prefWalk struct {
	pending  []*api.Task // ascending task ID
	group    map[string]*api.Task
	lastFail []C.uint32_t
	failed   error
	size     int // maxAssignments of the tree building: the size of every leaf's heap
}

// fillLeafStepwise: a leaf visit of a group whose cluster volumes count their users.  The leaf's heap is made once -- the k
// best nodes of (leaf AND volume set) as the tree building saw them: nothing has touched this leaf's nodes since, other
// leaves hold other nodes -- and every visit walks scheduleNTasksOnNodes over it with one engine question per step.  A
// visit after the first finds the heap collapsed (decision_tree.go:30-45): every node is put to the pipeline again, the
// ones that fail leave the heap for good, the rest are ordered by their rank of now.
func (s *Scheduler) fillLeafStepwise(ctx context.Context, n int, leaf *prefTree, w *prefWalk, decisions map[string]schedulingDecision, now time.Time) int {
	m := min(n, len(w.pending))
	if m <= 0 || w.failed != nil {
		return 0
	}
	sub, t := w.pending[:m], w.pending[0]
	if !leaf.built {
		cand, _, err := s.probeBest(t, leaf.leaf, w.size, now)
		if errors.Is(err, errUnsupported) {
			return 0
		} else if err != nil {
			w.failed = err
			return 0
		}
		leaf.cand, leaf.built = cand, true
	} else {
		ok, _, err := s.fitMany(t, leaf.cand, true, now)
		if err != nil {
			w.failed = err
			return 0
		}
		f := &VolumesFilter{vs: s.volumes}
		f.SetTask(t)
		var kept []rankKey
		for i, id := range leaf.cand {
			if ni, _ := s.nodeSet.nodeInfo(id); ok[i] == 1 && f.Check(&ni) {
				kept = append(kept, s.rankOf(id, t, now))
			}
		}
		sort.Slice(kept, func(i, j int) bool { return kept[i].less(kept[j]) })
		leaf.cand = leaf.cand[:0]
		for _, rk := range kept {
			leaf.cand = append(leaf.cand, rk.id)
		}
	}
	cnt := make([]C.uint32_t, C.PE_NUM_FILTERS)
	placed, _, err := s.fillStepwise(sub, leaf.cand, cnt, w.group, decisions, now)
	if err != nil {
		w.failed = err
	}
	w.pending = w.pending[placed:]
	if placed < m {
		w.lastFail = cnt
	}
	return placed
}

// fillLeaf: n tasks on one leaf = one engine group with leaf_cnt > 0 (the leaf's k best feasible nodes and
// scheduleNTasksOnNodes over them, on the device).  Why this is exact: DESIGN.md, "Placement preferences".
func (s *Scheduler) fillLeaf(ctx context.Context, n int, leaf *prefTree, w *prefWalk, decisions map[string]schedulingDecision, now time.Time) int {
	if s.gpu.leafStepwise {
		return s.fillLeafStepwise(ctx, n, leaf, w, decisions, now)
	}
	m := n
	if m > len(w.pending) {
		m = len(w.pending)
	}
	if m <= 0 || w.failed != nil {
		return 0
	}
	sub := w.pending[:m]
	var b tickBuf
	if err := s.encodeGroup(sub, &b, now); err != nil {
		return 0
	}
	g := &b.groups[len(b.groups)-1]
	if g.con_cnt == 0 {
		g.con_off = C.uint32_t(len(b.cons)) // the leaf terms follow the group's constraints
	}
	b.cons = append(b.cons, leaf.leaf...)
	g.leaf_cnt = C.uint32_t(len(leaf.leaf))
	if s.gpu.leafAlsoInVolumeSet { // a preference group with cluster mounts (scheduleVolumeGroup)
		b.cons = append(b.cons, C.pe_constraint{col: C.uint32_t(s.gpu.volCol), value: C.uint32_t(volMark)})
		g.leaf_cnt++
	}
	if w.failed = s.gpu.flushRows(&s.nodeSet); w.failed != nil {
		return 0
	}
	outNode := make([]C.uint32_t, m)
	outFail := make([]C.uint32_t, C.PE_NUM_FILTERS)
	tick := b.view()
	if rc := C.pe_schedule(s.gpu.h, &tick, ptr(outNode), ptr(outFail)); rc != C.PE_OK {
		w.failed = s.gpu.err("pe_schedule")
		return 0
	}
	placed := 0
	var left []*api.Task
	for i, t := range sub {
		if idx := uint32(outNode[i]); idx != C.PE_NONE {
			s.assign(t, s.gpu.order[idx], w.group, decisions)
			placed++
		} else {
			left = append(left, t)
		}
	}
	w.pending = append(left, w.pending[m:]...)
	if placed < m {
		w.lastFail = outFail
	}
	return placed
}

func (s *Scheduler) scheduleNTasksOnSubtreeGPU(ctx context.Context, n int, tree *prefTree, w *prefWalk, decisions map[string]schedulingDecision, now time.Time) int {
	if tree.next == nil {
		return s.fillLeaf(ctx, n, tree, w, decisions, now)
	}
	keys := make([]string, 0, len(tree.next))
	for k := range tree.next {
		keys = append(keys, k)
	}
	sort.Strings(keys) // canonical branch order (the reference ranges over a map, scheduler.go:796)
	tasksScheduled, tasksInUsableBranches := 0, tree.tasks
	noRoom := map[*prefTree]struct{}{}
	converging := true
	for tasksScheduled != n && len(noRoom) != len(tree.next) && converging {
		desired := (tasksInUsableBranches + n - tasksScheduled) / (len(tree.next) - len(noRoom))
		remainder := (tasksInUsableBranches + n - tasksScheduled) % (len(tree.next) - len(noRoom))
		converging = false
		for _, k := range keys {
			subtree := tree.next[k]
			if _, full := noRoom[subtree]; full {
				continue
			}
			if subtree.tasks < desired || (subtree.tasks == desired && remainder > 0) {
				converging = true
				toAssign := desired - subtree.tasks
				if remainder > 0 {
					toAssign++
				}
				res := s.scheduleNTasksOnSubtreeGPU(ctx, toAssign, subtree, w, decisions, now)
				if res < toAssign {
					noRoom[subtree] = struct{}{}
					tasksInUsableBranches -= subtree.tasks
				} else if remainder > 0 {
					remainder--
				}
				tasksScheduled += res
			}
		}
	}
	return tasksScheduled
}

func (s *Scheduler) schedulePreferenceGroup(ctx context.Context, group map[string]*api.Task, decisions map[string]schedulingDecision, now time.Time) {
	ts := sortedTasks(group)
	cols := s.preferenceLevels(ts[0])
	requeue := func(err error) {
		log.G(ctx).WithError(err).Error("placement engine")
		for _, t := range group {
			s.enqueue(t)
		}
		s.gpu.nodeSetChanged()
	}
	if err := s.gpu.flushRows(&s.nodeSet); err != nil { // (a new preference column is filled for every node)
		requeue(err)
		return
	}
	leafCap := len(s.gpu.order) + 1
	vals := make([]C.uint32_t, leafCap*len(cols))
	tasks := make([]C.uint32_t, leafCap)
	ccols := make([]C.uint32_t, len(cols))
	for i, c := range cols {
		ccols[i] = C.uint32_t(c)
	}
	var nLeaves C.uint32_t
	if rc := C.pe_pref_leaves(s.gpu.h, C.uint32_t(s.gpu.serviceID(ts[0].ServiceID)), ptr(ccols), C.uint32_t(len(cols)),
		C.uint32_t(leafCap), ptr(vals), ptr(tasks), &nLeaves); rc != C.PE_OK {
		requeue(s.gpu.err("pe_pref_leaves"))
		return
	}
	root := &prefTree{}
	for i := 0; i < int(nLeaves); i++ { // nodeset.go:59-101, one leaf's worth of nodes at a time
		tr := root
		var path []C.pe_constraint
		for l, col := range cols {
			tr.tasks += int(tasks[i])
			v := uint32(vals[i*len(cols)+l])
			if tr.next == nil {
				tr.next = map[string]*prefTree{}
			}
			nx := tr.next[s.gpu.prefStr[v]]
			if nx == nil {
				nx = &prefTree{}
				tr.next[s.gpu.prefStr[v]] = nx
			}
			tr = nx
			path = append(path, C.pe_constraint{col: C.uint32_t(col), value: C.uint32_t(v)})
		}
		tr.tasks += int(tasks[i])
		tr.leaf = path
	}
	w := &prefWalk{pending: ts, group: group, size: len(ts)}
	s.scheduleNTasksOnSubtreeGPU(ctx, len(ts), root, w, decisions, now)
	if w.failed != nil {
		requeue(w.failed)
		return
	}
	if len(group) != 0 {
		why := ""
		if w.lastFail != nil {
			why = explainCounters(w.lastFail)
		}
		s.noSuitableNodeWith(ctx, group, decisions, why)
	}
}

// ---- taskFitNode (scheduler.go:646-690) -----------------------------------------------------------------------------

func (s *Scheduler) taskFitNodeGPU(ctx context.Context, t *api.Task, nodeID string) *api.Task {
	if _, err := s.nodeSet.nodeInfo(nodeID); err != nil {
		return nil // node does not exist in set (it may have been deleted)
	}
	newT := *t
	var b tickBuf
	if err := s.encodeGroup([]*api.Task{t}, &b, time.Now()); err != nil {
		newT.Status.Err = err.Error()
		s.allTasks[t.ID] = &newT
		return &newT
	}
	if err := s.gpu.flushRows(&s.nodeSet); err != nil {
		log.G(ctx).WithError(err).Error("placement engine")
		return nil
	}
	idx := C.uint32_t(s.gpu.index[nodeID])
	var ok C.uint8_t
	outFail := make([]C.uint32_t, C.PE_NUM_FILTERS)
	tick := b.view()
	if rc := C.pe_fit(s.gpu.h, &tick, &idx, &ok, ptr(outFail)); rc != C.PE_OK || ok == 2 {
		return nil
	}
	if ok == 0 { // this node cannot accommodate this task
		newT.Status.Timestamp = ptypes.MustTimestampProto(time.Now())
		newT.Status.Err = explainCounters(outFail)
		s.allTasks[t.ID] = &newT
		return &newT
	}
	if hasClusterMounts(t) {
		// VolumesFilter is the last filter of the pipeline: it names the error only when every device filter passed; the
		// reservation pe_fit just made is taken back by uploading the row again from NodeInfo
		ni, _ := s.nodeSet.nodeInfo(nodeID)
		f := &VolumesFilter{vs: s.volumes}
		f.SetTask(t)
		verr := ""
		if !f.Check(&ni) {
			verr = f.Explain(1)
		} else if attachments, err := s.volumes.chooseTaskVolumes(t, &ni); err != nil { // scheduler.go:664-675
			verr = err.Error()
		} else {
			newT.Volumes = attachments
		}
		if verr != "" {
			s.gpu.nodeChanged(nodeID)
			newT.Status.Timestamp = ptypes.MustTimestampProto(time.Now())
			newT.Status.Err = verr
			s.allTasks[t.ID] = &newT
			return &newT
		}
	}
	newT.Status = api.TaskStatus{State: api.TaskStateAssigned, Timestamp: ptypes.MustTimestampProto(time.Now()),
		Message: "scheduler confirmed task can run on preassigned node"}
	s.allTasks[t.ID] = &newT
	if ni, err := s.nodeSet.nodeInfo(nodeID); err == nil && ni.addTask(&newT) {
		s.nodeSet.updateNode(ni)
	}
	return &newT
}

// ---- Pipeline.Explain (pipeline.go:84-103) from the engine's eight first-failing-filter counters ---------------------

func explainCounters(cnt []C.uint32_t) string {
	one := [8]string{"1 node not available for new tasks", "insufficient resources on 1 node", "missing plugin on 1 node",
		"scheduling constraints not satisfied on 1 node", "unsupported platform on 1 node", "host-mode port already in use on 1 node",
		"max replicas per node limit exceed", "cannot fulfill requested CSI volume mounts on 1 node"}
	many := [8]string{"%d nodes not available for new tasks", "insufficient resources on %d nodes", "missing plugin on %d nodes",
		"scheduling constraints not satisfied on %d nodes", "unsupported platform on %d nodes", "host-mode port already in use on %d nodes",
		"max replicas per node limit exceed", "cannot fulfill requested CSI volume mounts on %d nodes"}
	order := []int{0, 1, 2, 3, 4, 5, 6, 7}
	sort.SliceStable(order, func(i, j int) bool { return cnt[order[i]] > cnt[order[j]] }) // ties keep pipeline order
	var parts []string
	for _, f := range order {
		switch {
		case cnt[f] == 0:
		case cnt[f] == 1 || f == C.PE_F_MAXREPLICAS:
			parts = append(parts, one[f])
		default:
			parts = append(parts, fmt.Sprintf(many[f], uint32(cnt[f])))
		}
	}
	return strings.Join(parts, "; ")
}

// noSuitableNodeWith is noSuitableNode (scheduler.go:928-971) with the explanation passed in instead of read from
// s.pipeline.Explain(); the rest of that function is unchanged.
func (s *Scheduler) noSuitableNodeWith(ctx context.Context, group map[string]*api.Task, decisions map[string]schedulingDecision, explanation string) {
	for _, t := range group {
		var service *api.Service
		s.store.View(func(tx store.ReadTx) { service = store.GetService(tx, t.ServiceID) })
		if service == nil {
			continue
		}
		newT := *t
		newT.Status.Timestamp = ptypes.MustTimestampProto(time.Now())
		sv := service.SpecVersion
		if sv != nil && t.SpecVersion != nil && sv.Index > t.SpecVersion.Index {
			if t.Status.State == api.TaskStatePending && t.DesiredState >= api.TaskStateShutdown {
				newT.Status.State = api.TaskStateShutdown
				newT.Status.Err = ""
			}
		} else {
			if explanation != "" {
				newT.Status.Err = "no suitable node (" + explanation + ")"
			} else {
				newT.Status.Err = "no suitable node"
			}
			s.enqueue(&newT)
		}
		s.allTasks[t.ID] = &newT
		decisions[t.ID] = schedulingDecision{old: t, new: &newT}
	}
}
