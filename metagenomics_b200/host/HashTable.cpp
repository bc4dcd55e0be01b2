// HashTable.cpp -- see HashTable.h.
#include "HashTable.h"

HashTable::HashTable(void)
	: dataSet(NULL), hashTableSize(0), hashStringLength(0), numberOfHashCollision(0), context(NULL), ownsContext(true)
{
	const char *dev = getenv("OGB_DEVICE");
	ogbCheck(ogb_context_create(&context, dev ? atoi(dev) : 0), "HashTable");
}

HashTable::HashTable(ogb_context *ctx)
	: dataSet(NULL), hashTableSize(0), hashStringLength(0), numberOfHashCollision(0), context(ctx), ownsContext(false)
{
}

HashTable::~HashTable()
{
	for (map<string, vector<UINT64> *>::iterator it = lookupCache.begin(); it != lookupCache.end(); ++it) delete it->second;
	if (ownsContext && context) ogb_context_destroy(context);
}

// HashTable::insertDataset (HashTable.cpp:50-80): reads go to HBM once (K0 packs both strands), then
// K1 inserts the four prefix/suffix keys of every read.
bool HashTable::insertDataset(Dataset *d, UINT64 minOverlapLength)
{
	dataSet = d;
	hashStringLength = (UINT16)(minOverlapLength - 1);
	numberOfHashCollision = 0;
	ogbCheck(ogb_reads_upload_dataset(context, d->handle()), "HashTable::insertDataset");
	ogbCheck(ogb_hash_build(context, (uint32_t)minOverlapLength), "HashTable::insertDataset");
	hashTableSize = ogb_hash_table_size(context);
	return true;
}

// HashTable::getListOfReads (HashTable.cpp:202-221): never NULL, empty on a miss; entries are
// id | orientation<<62 in ascending (id, orientation) order.
vector<UINT64> *HashTable::getListOfReads(string subString)
{
	map<string, vector<UINT64> *>::iterator it = lookupCache.find(subString);
	if (it != lookupCache.end()) return it->second;
	vector<UINT64> *list = new vector<UINT64>;
	if (subString.size() == hashStringLength) {
		uint64_t offs[2] = {0, 0};
		vector<uint64_t> buf(64);
		int rc = ogb_hash_lookup(context, subString.c_str(), 1, buf.data(), buf.size(), offs);
		if (rc == OGB_E_CAPACITY) {
			buf.resize(offs[1]);
			rc = ogb_hash_lookup(context, subString.c_str(), 1, buf.data(), buf.size(), offs);
		}
		ogbCheck(rc, "HashTable::getListOfReads");
		for (uint64_t i = 0; i < offs[1]; i++) list->push_back(buf[i]);
	}
	lookupCache[subString] = list;
	return list;
}

// HashTable::hashFunction (HashTable.cpp:135-155), kept for interface parity only: the device
// table uses its own 64-bit mix (the hash value is unobservable in the result).
UINT64 HashTable::hashFunction(string subString)
{
	UINT64 first = 1, second = 1;
	for (size_t i = 0; i < subString.size(); i++) {
		UINT64 code = ((UINT64)subString[i] >> 1) & 3;
		if (i < 32) first = (first << 2) | code; else second = (second << 2) | code;
	}
	UINT64 p = hashTableSize ? hashTableSize : 1;
	return ((first % p) * (second % p)) % p;
}
