/* xmlmini/libxml/tree.h -- part of the libxml2-free reader of the drop-in build (see ../xml_mini.c).
 * Minimal stand-in for libxml2's DOM types (libxml2 is not installed in this image).  Only the
 * members the reference touches are present (state.c:260-411: doc->children, node->name,
 * node->children, node->next, xmlGetProp, xmlFree, xmlFreeDoc).  Implemented in ../xml_mini.c.
 */
#ifndef SBG_STUB_LIBXML_TREE_H
#define SBG_STUB_LIBXML_TREE_H

typedef unsigned char xmlChar;

typedef struct _xmlAttrMini {
  struct _xmlAttrMini *next;
  char *name;
  char *value;
} xmlAttrMini;

typedef struct _xmlNode {
  const xmlChar *name;
  struct _xmlNode *children;
  struct _xmlNode *next;
  struct _xmlNode *last_child; /* builder convenience */
  xmlAttrMini *attrs;
} xmlNode;
typedef xmlNode *xmlNodePtr;

typedef struct _xmlDoc {
  xmlNode *children;
} xmlDoc;
typedef xmlDoc *xmlDocPtr;

xmlChar *xmlGetProp(const xmlNode *node, const xmlChar *name);
void xmlFreeDoc(xmlDocPtr doc);
void sbg_xml_free(void *p);
#define xmlFree(p) sbg_xml_free((void *)(p))

#endif
